"""Which part of the path carries the TF32 logits error at config2?  Runs the product forward with the frustum-space
stages (view transformer) and the voxel-space stages (encoder / neck / head) in independently chosen math modes and
prints the error of every voxel-space stage against the reference golden (tests/golden/golden_config2.npz).
    python tools/parity_split_experiment.py [config2]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from util import build_model, full_inputs, golden_full, sample_stage, stage_error   # noqa: E402
from stereoscene_b200 import ops                                                       # noqa: E402


def run(workload, policy):
    meta, gold = golden_full(workload)
    model, mc = build_model(workload, meta["seed"], device="cuda")
    xl, xr, left, right, calib = full_inputs(meta, device="cuda")
    vt = model.img_view_transformer
    vt.stage_outputs = {}
    ops.set_math_policy(policy)
    with torch.no_grad():
        out = model.forward_features(xl, xr, left, right, calib, occ_size=meta["occ_size"])
    torch.cuda.synchronize()
    ops.set_math_policy(None)
    st = vt.stage_outputs
    got = dict(stereo_prob=st["stereo_prob"], lss_prob=st["lss_prob"], depth_prob=out["depth"], bev_feat=st["bev_feat"],
               enc0=st["enc0"], neck=st["neck"], logits=out["logits_lowres"], logits_up=out["output_voxels"])
    name = ",".join(f"{g}" for g, m in policy.items() if m != ops.SS_MATH_TF32) or "none"
    print(f"compensated: {name:24s} " + "  ".join(
        f"{k}:{stage_error(sample_stage(v, meta, k), gold[k], meta['stats'][k])['max']:.2e}/"
        f"{stage_error(sample_stage(v, meta, k), gold[k], meta['stats'][k])['rms']:.2e}" for k, v in got.items()), flush=True)


SUB = ("depthnet.trunk", "depthnet.blocks", "depthnet.aspp", "depthnet.dcn", "mie.redir1", "mie.hourglass", "mie.ca3d", "mie.redir2")

if __name__ == "__main__":
    import itertools
    wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
    what = sys.argv[2] if len(sys.argv) > 2 else "groups"
    X3, T = ops.SS_MATH_TF32X3, ops.SS_MATH_TF32
    if what == "groups":
        groups = ("stereo", "depthnet", "mie")
        for mask in itertools.product((0, 1), repeat=3):
            run(wl, {g: (X3 if m else T) for g, m in zip(groups, mask)})
        run(wl, {g: X3 for g in groups + ("voxel",)})
    elif what == "leave-one-out":          # depth_net + MIE compensated except one sub-stage (printed names = what IS compensated)
        for drop in SUB:
            run(wl, {g: (T if g == drop else X3) for g in SUB})
    else:                                   # explicit list of compensated sub-stages: a,b,c
        keep = what.split(",")
        run(wl, {g: (X3 if g in keep else T) for g in SUB})
