"""Row N2: time of the image encoder (EfficientNet-B7 + SECONDFPN) on one 384x1280 stereo pair, eager and as a CUDA graph,
per math policy, with the kernel census of one forward.
    python tools/image_encoder_timing.py [policy ...]      (default: mixed tf32)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from stereoscene_b200 import cabi, ops, synth
from util import build_image_encoder

dev = torch.device("cuda", 0)
enc = build_image_encoder(0, dev)
left, right = synth.stereo_images(1, (384, 1280), seed=0, device=dev)
img = torch.cat([left, right], 0).flatten(0, 1)


def fwd():
    ops.arena(dev).reset()
    return enc["img_neck"].forward_vol(enc["img_backbone"].forward_vol(img))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


if os.environ.get("PROFILE_ONE"):           # under ncu --profile-from-start off: one warm forward in the given policy
    ops.set_math_policy(os.environ["PROFILE_ONE"])
    with torch.no_grad():
        fwd(); fwd()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fwd()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    sys.exit(0)

for pol in (sys.argv[1:] or ["mixed", "tf32"]):
    ops.set_math_policy(pol)
    with torch.no_grad():
        fwd()
        c0 = cabi.kernel_census()
        fwd()
        c1 = cabi.kernel_census()
        eager = timed(fwd)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            fwd()
        graph = timed(g.replay)
    census = {k: c1[k] - c0.get(k, 0) for k in c1 if c1[k] - c0.get(k, 0)}
    print(f"policy {pol}: eager {eager:.3f} ms, graph {graph:.3f} ms, {sum(census.values())} launches: {census}", flush=True)
    del g
