"""TEST INFRASTRUCTURE ONLY -- golden fixtures of the 2-D image encoder written by
the reference's own efficientnet.py (loaded unmodified through oracle/ref_loader.py, on the restated mmcv / mmdet
bricks and the restated mmdet3d SECONDFPN that ref_loader documents) on seeded synthetic stereo images.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_image
  tests/golden/golden_image_tiny.npz   64x128 images, one stereo pair: every backbone level and the neck output, complete
  tests/golden/golden_image_full.npz   384x1280 (BASELINE.json configs[1..4] input size): strided samples + full-tensor
                                       max|.| / RMS of the same tensors (slices in the json, re-applied by the tests)
  tests/golden/state_dict_spec.json    gains the img_backbone.* / img_neck.* keys (checkpoint contract of the encoder)
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_loader as R                     # noqa: E402
from stereoscene_b200 import presets, synth            # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
OUT_INDICES = (2, 3, 4, 5, 6)
S = lambda a=0, s=1: (a, None, s)       # noqa: E731
SAMPLERS = {          # [B*N, C, H, W]
    "img_level2": [S(), S(1, 4), S(0, 4), S(1, 4)],
    "img_level3": [S(), S(0, 4), S(1, 2), S(0, 4)],
    "img_level4": [S(), S(2, 8), S(0, 2), S(1, 2)],
    "img_level5": [S(), S(3, 16), S(), S(0, 2)],
    "img_level6": [S(), S(5, 32), S(), S(1, 2)],
    "img_feat": [S(), S(0, 8), S(1, 2), S(0, 4)],
}
CASES = {"tiny": dict(input_size=(64, 128), seed=5), "full": dict(input_size=(384, 1280), seed=0)}


def build_reference_encoder():
    """img_backbone / img_neck of stereoscene.py:59-74, built from the reference class and the shipped config."""
    cfg = presets.shipped_config()["model"]
    bb = dict(cfg["img_backbone"]); bb.pop("type"); bb.pop("init_cfg", None)
    nk = dict(cfg["img_neck"]); nk.pop("type")
    net = R.efficientnet_module().CustomEfficientNet(**bb)
    neck = R.SECONDFPN(**nk)
    model = torch.nn.ModuleDict(dict(img_backbone=net, img_neck=neck))
    for m in model.modules():
        m.training = False              # the reference's train() override returns None, so .eval() cannot be chained
    return model


@torch.no_grad()
def run_reference(model, img):
    """DET:42-59 on [B,N,3,H,W]: returns dict(img_level{i}, img_feat [B*N,640,H/8,W/8])."""
    B, N, Cc, H, W = img.shape
    levels = model["img_backbone"](img.view(B * N, Cc, H, W))
    x = model["img_neck"](levels)[0]
    st = {f"img_level{i}": t for i, t in zip(OUT_INDICES, levels)}
    st["img_feat"] = x
    return st


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    model = build_reference_encoder()
    spec_path = os.path.join(GOLDEN_DIR, "state_dict_spec.json")
    with open(spec_path) as f:
        spec = json.load(f)
    spec = {k: v for k, v in spec.items() if not k.startswith(("img_backbone.", "img_neck."))}
    spec.update({k: list(v.shape) for k, v in model.state_dict().items()})
    with open(spec_path, "w") as f:
        json.dump(spec, f, indent=0, sort_keys=True)
    for name, case in CASES.items():
        synth.randomize_weights_(model, case["seed"])
        left, right = synth.stereo_images(1, case["input_size"], seed=case["seed"])
        img = torch.cat([left, right], 0)                   # DET:94 torch.cat([img[0], img2[0]], 0)
        t0 = time.time()
        st = run_reference(model, img)
        dt = time.time() - t0
        arrays, stats = {}, {}
        for k, t in st.items():
            t = t.float()
            arrays[k] = t.numpy() if name == "tiny" else t[tuple(slice(*s) for s in SAMPLERS[k])].contiguous().numpy()
            stats[k] = dict(shape=list(t.shape), absmax=float(t.abs().max()), rms=float(t.double().pow(2).mean().sqrt()))
        path = os.path.join(GOLDEN_DIR, f"golden_image_{name}.npz")
        np.savez_compressed(path, **arrays)
        meta = dict(case=name, seed=case["seed"], batch=1, input_size=list(case["input_size"]),
                    samplers=({} if name == "tiny" else SAMPLERS), stats=stats, reference_forward_seconds=round(dt, 2))
        with open(os.path.join(GOLDEN_DIR, f"golden_image_{name}.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(f"{name}: reference forward {dt:.1f} s, {os.path.getsize(path)/1e6:.2f} MB")
        for k in st:
            print(f"  {k:12s} {tuple(stats[k]['shape'])} -> {arrays[k].shape} absmax {stats[k]['absmax']:.3e} rms {stats[k]['rms']:.3e}")


if __name__ == "__main__":
    main()
