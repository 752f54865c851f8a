#!/usr/bin/env python
"""Bisect which part of depth_net disagrees under CUDA-graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stereoscene_b200 import presets, synth
from stereoscene_b200.plugin.view_transformer import _group_norm_wide

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))

dev = torch.device("cuda", 0)
model, mc = presets.build("config0")
synth.randomize_weights_(model, 0)
model = model.to(dev).eval()
dn = model.img_view_transformer.depth_net
x = torch.randn(1, 640, 16, 32, device=dev)
m = torch.randn(1, 30, device=dev)

def run():
    o = {}
    mm = dn.bn(m)
    o["bn"] = mm
    y = dn.reduce_conv[0](x); o["conv0"] = y
    y = torch.relu_(_group_norm_wide(y, dn.reduce_conv[1])); o["gn"] = y
    c = dn.context_se(y, dn.context_mlp(mm)[..., None, None]); o["ctx_se"] = c
    c = dn.context_conv(c); o["ctx"] = c
    d = dn.depth_se(y, dn.depth_mlp(mm)[..., None, None]); o["d_se"] = d
    for i in range(3):
        d = dn.depth_conv[i](d); o[f"bb{i}"] = d
    d = dn.depth_conv[3](d); o["aspp"] = d
    d = dn.depth_conv[4](d); o["dcn"] = d
    d = dn.depth_conv[5](d); o["out"] = d
    return o

with torch.no_grad():
    run(); e = {k: v.clone() for k, v in run().items()}
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        og = run()
    g.replay(); torch.cuda.synchronize()
    print({k: f"{rel(og[k], e[k]):.1e}" for k in e})
