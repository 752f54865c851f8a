#!/usr/bin/env python
"""Compare an eager forward with a CUDA-graph replay of the same forward, stage by stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stereoscene_b200 import ops, presets, synth

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))

dev = torch.device("cuda", 0)
wl = sys.argv[1] if len(sys.argv) > 1 else "config0"
model, mc = presets.build(wl)
synth.randomize_weights_(model, 0)
model = model.to(dev).eval()
xl, xr = synth.stereo_features(1, mc["input_size"], 8, seed=0, device=dev)
left, right, calib = synth.kitti_calibration(1, mc["input_size"], device=dev)
vt = model.img_view_transformer

def fwd():
    vt.stage_outputs = {}
    out = model.forward_features(xl, xr, left, right, calib, occ_size=mc["occ_size"], want_labels=True)
    st = dict(vt.stage_outputs); vt.stage_outputs = None
    return out, st

with torch.no_grad():
    fwd(); out_e, st_e = fwd()
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in st_e.items() if torch.is_tensor(v)}
    eager.update({k: v.clone() for k, v in out_e.items() if torch.is_tensor(v)})
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fwd()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_g, st_g = fwd()
    for rep in range(2):
        g.replay(); torch.cuda.synchronize()
        got = {k: v for k, v in st_g.items() if torch.is_tensor(v)}
        got.update({k: v for k, v in out_g.items() if torch.is_tensor(v)})
        print("replay", rep, {k: f"{rel(got[k].float(), eager[k].float()):.2e}" for k in eager if k in got})
