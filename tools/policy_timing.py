"""Step time (CUDA-graph replay) and logits parity of arbitrary math policies at config2.
    python tools/policy_timing.py 'name=group:mode,group:mode;name2=...'   (modes: tf32 f16 x3)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from stereoscene_b200 import ops, presets, synth

M = {"tf32": ops.SS_MATH_TF32, "f16": ops.SS_MATH_F16, "x3": ops.SS_MATH_TF32X3}
dev = torch.device("cuda", 0)
WL = os.environ.get("WORKLOAD", "config2")
SEED = json.load(open(os.path.join(ROOT, "tests", "golden", f"golden_{WL}.json")))["seed"]
model, mc = presets.build(WL)
synth.randomize_weights_(model, SEED)
model = model.to(dev).eval()
xl, xr = synth.stereo_features(1, mc["input_size"], 8, seed=SEED, device=dev)
left, right, calib = synth.kitti_calibration(1, mc["input_size"], device=dev)
occ = mc["occ_size"]


def fwd():
    return model.forward_features(xl, xr, left, right, calib, occ_size=occ, want_labels=True)


specs = sys.argv[1].split(";") if len(sys.argv) > 1 else ["mixed=depthnet:x3,mie:x3,mie.ca3d:tf32"]
for spec in specs:
    name, _, body = spec.partition("=")
    pol = {g: M[m] for g, m in (kv.split(":") for kv in body.split(",") if kv)}
    ops.set_math_policy(pol)
    with torch.no_grad():
        for _ in range(2):
            o = fwd()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fwd()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    par = bench.golden_parity(WL, o, SEED)
    print(f"{name:28s} {a.elapsed_time(b) / 20:7.3f} ms   logits {par['logits']['max_rel']:.2e}/{par['logits']['rms_rel']:.2e}  "
          f"logits_up {par['logits_up']['max_rel']:.2e}/{par['logits_up']['rms_rel']:.2e}", flush=True)
    del g
