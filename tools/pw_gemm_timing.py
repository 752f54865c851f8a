"""Timing of single pointwise (1x1) convolutions of the image encoder's shapes on the tcgen05 kernels (L2-warm, CUDA-graph replay of 50 calls).
    python tools/pw_gemm_timing.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
from stereoscene_b200 import ops
from stereoscene_b200.ops import Vol

dev = torch.device("cuda", 0)
MODES = {"tf32": ops.SS_MATH_TF32, "x3": ops.SS_MATH_TF32X3}
ACTS = {"none": ops.SS_ACT_NONE, "swish": ops.SS_ACT_SWISH, "relu": ops.SS_ACT_RELU}


def run(N, H, W, cin, cout, mode, act, gate=False, acc=False, iters=50):
    conv = nn.Conv2d(cin, cout, 1, bias=True).to(dev)
    x = torch.randn(N, 1, H, W, cin, device=dev)
    out = torch.zeros(N, 1, H, W, cout, device=dev)
    v = Vol(x, torch.rand(N, cin, device=dev), torch.zeros(N, cin, device=dev)) if gate else Vol(x)
    kw = dict(accumulate=True) if acc else {}
    f = lambda: ops.conv(v, conv, out=out, out_act=ACTS[act], math_mode=MODES[mode], **kw)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()                 # a graph of back-to-back calls: no host launch cost in the number
    with torch.cuda.graph(g):
        for _ in range(iters):
            f()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / iters * 1e3
    gf = 2.0 * N * H * W * cin * cout / 1e9
    print(f"{N}x{H}x{W} {cin:5d}->{cout:5d} {mode:5s} act={act:5s} gate={int(gate)} acc={int(acc)}: {us:8.1f} us  {gf / us * 1e3:8.1f} TFLOP/s", flush=True)


SHAPES = [(2, 96, 320, 64, 288), (2, 96, 320, 288, 48), (2, 24, 80, 160, 960), (2, 24, 80, 960, 160), (2, 24, 80, 224, 1344),
          (2, 24, 80, 1344, 224), (2, 12, 40, 384, 2304), (2, 12, 40, 2304, 384), (2, 12, 40, 640, 3840), (2, 12, 40, 3840, 640)]
if __name__ == "__main__":
    for (N, H, W, ci, co) in SHAPES:
        expand = co > ci
        for mode in ("tf32", "x3"):
            run(N, H, W, ci, co, mode, "swish" if expand else "none", gate=not expand)
        if expand:
            run(N, H, W, ci, co, "x3", "none")
            run(N, H, W, ci, min(co, 256), "x3", "swish")
