#!/usr/bin/env python
"""End-to-end use of the public API on synthetic stereo features (needs a B200):

  config file (the reference's stereoscene.py, unchanged, or the packaged copy of its model dict)
    -> registry-built BEVDepthOccupancy (this repo's modules under the reference's names)
    -> VolumetricEngine (CUDA graph + overlapped host copies) over a stream of stereo pairs
    -> SSCMetrics (one confusion-matrix kernel per sample) and SemanticKITTI .label files.

  python examples/run_inference.py [--pairs 8] [--workload config2] [--out /tmp/ssc_out] [--config path/to/stereoscene.py]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--config", default=None, help="the reference's stereoscene.py (loaded unchanged); default: packaged model dict")
    ap.add_argument("--out", default=None, help="write SemanticKITTI .label predictions under this folder")
    a = ap.parse_args()

    from stereoscene_b200 import presets, semkitti_io, synth
    from stereoscene_b200.plugin import SSCMetrics
    from stereoscene_b200.runtime import VolumetricEngine

    dev = torch.device("cuda", 0)
    if a.config:
        import projects.mmdet3d_plugin  # noqa: F401  (this repo's drop-in package: registers the modules)
        from stereoscene_b200.config import Config
        from stereoscene_b200.registry import build_model
        cfg = Config.fromfile(a.config)
        model = build_model(cfg.model, train_cfg=cfg.get("train_cfg"), test_cfg=cfg.get("test_cfg")).eval()
        occ_size, input_size = list(cfg.occ_size), tuple(cfg.data_config["input_size"])
    else:
        model, mc = presets.build(a.workload)
        occ_size, input_size = mc["occ_size"], mc["input_size"]
    synth.randomize_weights_(model, 0)                      # no checkpoint offline: seeded weights
    model = model.to(dev)

    left, right, calib = synth.kitti_calibration(1, input_size, device=dev)
    pairs = [synth.stereo_features(1, input_size, 8, seed=i, pin=True) for i in range(a.pairs)]
    eng = VolumetricEngine(model, left, right, calib, occ_size, tuple(pairs[0][0].shape), device=dev)

    metric = SSCMetrics().to(dev)
    gen = torch.Generator().manual_seed(0)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i, labels in enumerate(eng.stream(pairs)):          # uint8 [1, X, Y, Z] on the host, in input order
        target = torch.randint(0, 20, labels.shape, generator=gen, dtype=torch.uint8)      # synthetic ground truth
        target[torch.rand(labels.shape, generator=gen) < 0.1] = 255
        metric.update(labels.to(dev), target.to(dev))            # blocking copies: the yielded pinned buffer is reused two pairs later
        if a.out:
            semkitti_io.save_output_semantic_kitti(labels[0], a.out, "08", f"{i:06d}")
    stop.record()
    torch.cuda.synchronize()
    res = metric.compute()
    n_vox = a.pairs * occ_size[0] * occ_size[1] * occ_size[2]
    print(f"{a.pairs} pairs, {n_vox / (start.elapsed_time(stop) * 1e-3):.3e} voxels/s incl. scoring"
          f"{' and label files' if a.out else ''}; SC IoU {res['iou']:.4f}, SSC mIoU {res['iou_ssc_mean']:.4f} (random targets)")


if __name__ == "__main__":
    main()
