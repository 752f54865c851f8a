"""Shared helpers for the parity tests."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from stereoscene_b200 import presets, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b) -> float:
    """max|a-b| / max|b| on float64 numpy views."""
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def golden_tiny():
    with open(os.path.join(GOLDEN, "golden_tiny.json")) as f:
        cfg = json.load(f)
    return cfg, np.load(os.path.join(GOLDEN, "golden_tiny.npz"))


def build_model(workload: str, seed: int, device="cpu"):
    """Our detector for a workload with the seeded key-addressed weights."""
    model, mc = presets.build(workload)
    synth.randomize_weights_(model, seed)
    return model.to(device).eval(), mc


def tiny_inputs(cfg, device="cpu"):
    B = cfg["batch"]
    xl, xr = synth.stereo_features(B, tuple(cfg["input_size"]), cfg["downsample"], seed=cfg["seed"])
    left, right, calib = synth.kitti_calibration(B, tuple(cfg["input_size"]))
    calib = calib * torch.tensor(cfg.get("calib_scale", [1.0] * B)).view(B, 1)
    mv = lambda d: {k: v.to(device) for k, v in d.items()}   # noqa: E731
    return xl.to(device), xr.to(device), mv(left), mv(right), calib.to(device)


def cpu_state_dict(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


# ---- full-size fixtures (tests/golden/golden_config{1,2}.npz, written by oracle/make_golden_full.py) -------------
def golden_full(workload: str):
    with open(os.path.join(GOLDEN, f"golden_{workload}.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLDEN, f"golden_{workload}.npz"))


def full_inputs(meta, device="cpu"):
    """The seeded synthetic pair + KITTI calibration the fixture was generated from."""
    B = meta["batch"]
    xl, xr = synth.stereo_features(B, tuple(meta["input_size"]), 8, seed=meta["seed"])
    left, right, calib = synth.kitti_calibration(B, tuple(meta["input_size"]))
    mv = lambda d: {k: v.to(device) for k, v in d.items()}   # noqa: E731
    return xl.to(device), xr.to(device), mv(left), mv(right), calib.to(device)


def sample_stage(t, meta, key):
    """Re-apply the fixture's strided sampler to a full stage tensor (logical reference layout)."""
    sl = tuple(slice(*s) for s in meta["samplers"][key])
    return t[sl]


def stage_error(got_sample, want_sample, stat) -> dict:
    """Errors of a sampled stage against the reference's sample, normalised by FULL-tensor statistics of the
    reference (absmax, rms from the fixture's json): ``max`` = max|d| / absmax (the rel-to-max reading of the north
    star's "1e-3 relative"), ``rms`` = rms(d) / rms(ref) (relative L2 error, the stricter reading)."""
    a = got_sample.detach().cpu().double().numpy() if torch.is_tensor(got_sample) else np.asarray(got_sample, np.float64)
    b = np.asarray(want_sample, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = a - b
    return dict(max=float(np.abs(d).max() / stat["absmax"]), rms=float(np.sqrt((d * d).mean()) / stat["rms"]))


# ---- 2-D image encoder fixtures (tests/golden/golden_image_{tiny,full}.npz, written by oracle/make_golden_image.py) ----
def golden_image(case: str):
    with open(os.path.join(GOLDEN, f"golden_image_{case}.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLDEN, f"golden_image_{case}.npz"))


def image_inputs(meta, device="cpu"):
    """The seeded stereo pair of the fixture as the reference feeds it: torch.cat([left, right], 0) -> [2,1,3,H,W]."""
    left, right = synth.stereo_images(meta["batch"], tuple(meta["input_size"]), seed=meta["seed"])
    return torch.cat([left, right], 0).to(device)


def build_image_encoder(seed: int, device="cpu"):
    """Our CustomEfficientNet-B7 + SECONDFPN from the shipped config, with the seeded key-addressed weights under the
    detector's prefixes (img_backbone. / img_neck.)."""
    from stereoscene_b200 import plugin  # noqa: F401
    from stereoscene_b200.registry import build_backbone, build_neck
    cfg = presets.model_config("config2", image_encoder=True)["model"]
    enc = torch.nn.ModuleDict(dict(img_backbone=build_backbone(cfg["img_backbone"]), img_neck=build_neck(cfg["img_neck"])))
    synth.randomize_weights_(enc, seed)
    return enc.to(device).eval()
