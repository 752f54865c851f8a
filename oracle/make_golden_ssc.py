"""TEST INFRASTRUCTURE ONLY -- golden vectors for the SSC scores (row N4) from the reference's own
``SSCMetrics`` (projects/mmdet3d_plugin/utils/ssc_metric.py, imported unmodified through
oracle/ref_loader.py; torchmetrics' ``Metric`` base is a stand-in that only registers the buffers).

Run in the build container:   python -m oracle.make_golden_ssc   -> tests/golden/golden_ssc.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_loader as R                     # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cases(seed=5):
    g = torch.Generator().manual_seed(seed)
    out = []
    for ci, (shape, use_ne, use_ns) in enumerate((((2, 16, 16, 4), False, False), ((1, 12, 10, 6), True, False),
                                                   ((3, 8, 8, 8), True, True))):
        n = int(np.prod(shape))
        pred = torch.randint(0, 20, shape, generator=g)
        true = torch.randint(0, 20, shape, generator=g)
        true[torch.rand(shape, generator=g) < 0.35] = 0                      # plenty of empty space
        pred[torch.rand(shape, generator=g) < 0.30] = 0
        agree = torch.rand(shape, generator=g) < 0.4
        pred[agree] = true[agree]
        true[torch.rand(shape, generator=g) < 0.15] = 255                    # ignored voxels
        ne = (torch.rand(shape, generator=g) < 0.8) if use_ne else None
        ns = (torch.rand(shape, generator=g) < 0.7) if use_ns else None
        out.append((f"c{ci}", pred, true, ne, ns))
    return out


def main():
    mod = R._imp("projects.mmdet3d_plugin.utils.ssc_metric")
    arrays = {}
    metric = mod.SSCMetrics()
    for name, pred, true, ne, ns in cases():
        m1 = mod.SSCMetrics()
        m1.update(pred.clone(), true.clone(), None if ne is None else ne.clone(), None if ns is None else ns.clone())
        arrays[name + "_pred"] = pred.numpy().astype(np.uint8)
        arrays[name + "_true"] = true.numpy().astype(np.uint8)
        if ne is not None:
            arrays[name + "_nonempty"] = ne.numpy()
        if ns is not None:
            arrays[name + "_nonsurface"] = ns.numpy()
        arrays[name + "_completion"] = np.array([float(m1.completion_tp), float(m1.completion_fp), float(m1.completion_fn)])
        arrays[name + "_tps"], arrays[name + "_fps"], arrays[name + "_fns"] = m1.tps.numpy(), m1.fps.numpy(), m1.fns.numpy()
        metric.update(pred.clone(), true.clone(), None if ne is None else ne.clone(), None if ns is None else ns.clone())
    res = metric.compute()                                                    # accumulated over the three cases
    arrays["all_iou_ssc"] = res["iou_ssc"].numpy()
    arrays["all_scalars"] = np.array([float(res["precision"]), float(res["recall"]), float(res["iou"]), float(res["iou_ssc_mean"])])
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "golden_ssc.npz"), **arrays)
    print("wrote golden_ssc.npz:", {k: v.shape for k, v in arrays.items() if k.startswith("all") or k.endswith("tps")})


if __name__ == "__main__":
    main()
