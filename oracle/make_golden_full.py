"""TEST INFRASTRUCTURE ONLY -- full-size golden fixtures (BASELINE.json configs[1] and configs[2]) written by the
reference's own, unmodified module code (oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_full
The fixtures are committed; the GPU box (which has no reference tree) only reads them.

A full-size forward has ~10^8 stage values, so the fixtures hold *strided samples* of every stage boundary (the
slices are recorded in the json next to the npz and re-applied by the tests), the full-tensor max|.| and RMS of each
stage (so an error can be quoted relative to the whole tensor, not to the sample), and for the integer index path the
complete kept mask (bit-packed) plus the complete linear voxel ids (int32, -1 = dropped).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import make_golden as MG                   # noqa: E402
from oracle import restatement as O                    # noqa: E402
from stereoscene_b200 import presets, synth            # noqa: E402

GOLDEN_DIR = MG.GOLDEN_DIR

# stage -> slice per axis (start, stop, step); missing trailing axes are taken whole
S = lambda a=0, s=1: (a, None, s)       # noqa: E731
SAMPLERS = {
    "stereo_fea": [S(), S(0, 4), S(1, 2), S(1, 4)],
    "gwc_warp": [S(), S(1, 4), S(0, 4), S(1, 4), S(2, 4)],
    "stereo_prob": [S(), S(0, 2), S(1, 2), S(0, 2)],
    "lss_prob": [S(), S(1, 2), S(0, 2), S(1, 2)],
    "depth_net": [S(), S(0, 4), S(1, 2), S(0, 4)],
    "bri_lss2stereo": [S(), S(), S(0, 2), S(0, 2), S(1, 2)],
    "bri_stereo2lss": [S(), S(), S(1, 2), S(1, 2), S(0, 2)],
    "mie_hourglass": [S(), S(2, 4), S(1, 4), S(0, 4), S(1, 4)],
    "mie_ca3d": [S(), S(3, 4), S(2, 4), S(1, 4), S(0, 4)],
    "depth_prob": [S(), S(0, 2), S(0, 2), S(1, 2)],
    "bev_feat": [S(), S(0, 8), S(1, 4), S(2, 4), S(0, 2)],
    "enc0": [S(), S(3, 8), S(0, 4), S(1, 4), S(1, 2)],
    "enc1": [S(), S(5, 8), S(0, 2), S(1, 2), S(0, 2)],
    "enc2": [S(), S(7, 8), S(0, 2), S(1, 2), S()],
    "neck": [S(), S(0, 16), S(2, 4), S(0, 4), S(1, 2)],
    "logits": [S(), S(), S(1, 4), S(2, 4), S(0, 2)],
    "logits_up": [S(), S(), S(3, 8), S(5, 8), S(1, 4)],
}
FRUSTUM_KEYS = ("stereo_fea", "gwc_warp", "stereo_prob", "lss_prob", "depth_net", "bri_lss2stereo", "bri_stereo2lss",
                "mie_hourglass", "mie_ca3d")
VOXEL_KEYS = ("depth_prob", "bev_feat", "enc0", "enc1", "enc2", "neck", "logits", "logits_up")

# workload -> (seed, which stages are stored).  config1 and config2 share the frustum geometry (SURVEY.md section 8d:
# frustum-space work is independent of the grid), so config1 keeps only the grid-dependent stages and uses another seed.
CASES = {"config2": (0, FRUSTUM_KEYS + VOXEL_KEYS), "config1": (1, VOXEL_KEYS)}


def sample(t: torch.Tensor, key: str) -> torch.Tensor:
    sl = tuple(slice(*s) for s in SAMPLERS[key])
    return t[sl].contiguous()


def reference_cfg(workload: str, seed: int) -> dict:
    mc = presets.model_config(workload)
    vt = mc["model"]["img_view_transformer"]
    return dict(batch=1, input_size=tuple(mc["input_size"]), downsample=8, grid_config=vt["grid_config"],
                occ_size=mc["occ_size"], seed=seed)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    for workload, (seed, keys) in CASES.items():
        cfg = reference_cfg(workload, seed)
        model = MG.build_reference_model(cfg)
        synth.randomize_weights_(model, seed)
        xl, xr, left, right, calib = MG.synthetic_inputs(cfg)
        t0 = time.time()
        st = MG.run_reference(model, cfg, xl, xr, left, right, calib)
        dt = time.time() - t0
        st["neck"] = st["neck"]
        arrays, stats = {}, {}
        for k in keys:
            t = st[k].float()
            arrays[k] = sample(t, k).numpy()
            stats[k] = dict(shape=list(t.shape), absmax=float(t.abs().max()), rms=float(t.double().pow(2).mean().sqrt()))
        # integer index path: the reference's own voxel_pooling arithmetic (VT:441-449) on the reference's geometry
        gc = cfg["grid_config"]
        dx, bx, nx = O.gen_dx_bx(gc["xbound"], gc["ybound"], gc["zbound"])
        idx, kept = O.voxel_indices(st["geom"], dx, bx, nx)
        n = [int(v) for v in nx.tolist()]
        lin = (idx[:, 0] * n[1] + idx[:, 1]) * n[2] + idx[:, 2]
        lin = torch.where(kept, lin, torch.full_like(lin, -1)).to(torch.int32)
        arrays["kept_bits"] = np.packbits(kept.numpy().astype(np.uint8))
        arrays["voxel_lin"] = lin.numpy()
        arrays["labels_up_sample"] = sample(st["logits_up"], "logits_up").argmax(1).numpy().astype(np.uint8)
        arrays["calib"] = calib.numpy()
        path = os.path.join(GOLDEN_DIR, f"golden_{workload}.npz")
        np.savez_compressed(path, **arrays)
        meta = dict(workload=workload, seed=seed, batch=1, input_size=list(cfg["input_size"]), occ_size=cfg["occ_size"],
                    grid_config=gc, samplers={k: SAMPLERS[k] for k in keys}, stats=stats,
                    kept_points=int(kept.sum()), points=int(kept.numel()), reference_forward_seconds=round(dt, 2),
                    threads=torch.get_num_threads())
        with open(os.path.join(GOLDEN_DIR, f"golden_{workload}.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(f"{workload}: reference forward {dt:.1f} s, {len(arrays)} arrays, {os.path.getsize(path)/1e6:.2f} MB, "
              f"kept {int(kept.sum())}/{kept.numel()} points")
        for k in keys:
            print(f"  {k:16s} {tuple(stats[k]['shape'])} -> {arrays[k].shape}  absmax {stats[k]['absmax']:.4e} rms {stats[k]['rms']:.4e}")
        del model, st


if __name__ == "__main__":
    main()
