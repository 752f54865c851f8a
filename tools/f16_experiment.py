"""Per-kernel timing of the fp16-operand single pass (SS_MATH_F16) against TF32 and the compensated mode on the layers that
dominate the step (CUDA events, 20 launches each)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stereoscene_b200 import ops, presets, synth
from stereoscene_b200.ops import Vol

dev = torch.device("cuda", 0)
model, mc = presets.build("config2")
synth.randomize_weights_(model, 0)
model = model.to(dev).eval()
vt = model.img_view_transformer
nx = [int(round(float(v))) for v in vt.nx.detach().cpu()]
D, H, W = vt.D, vt.frustum.shape[1], vt.frustum.shape[2]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


cases = []
head = model.pts_bbox_head.occ_convs[0][0]
x = torch.randn((1, nx[0], nx[1], nx[2], 384), device=dev)
sc, sh = torch.rand((1, 384), device=dev) + 0.5, torch.randn((1, 384), device=dev) * 0.1
cases.append(("head 384->192 k3 (pending)", lambda m: ops.conv(Vol(x, sc, sh, ops.SS_ACT_RELU), head, want_stats=True, math_mode=m)))
c = model.img_bev_encoder_backbone.layers[0][0].conv1
xe = torch.randn((1, nx[0], nx[1], nx[2], 128), device=dev)
se, he = torch.rand((1, 128), device=dev) + 0.5, torch.randn((1, 128), device=dev) * 0.1
cases.append(("enc 128->128 k3 (pending)", lambda m: ops.conv(Vol(xe, se, he, ops.SS_ACT_RELU), c, want_stats=True, math_mode=m)))
cases.append(("enc 128->128 k3 (plain)", lambda m: ops.conv(Vol(xe), c, want_stats=True, math_mode=m)))
c32 = vt.stereo_volume_net.dres0[0][0]
xv = torch.randn((1, D, H, W, 32), device=dev)
s3, h3 = torch.rand((1, 32), device=dev) + 0.5, torch.randn((1, 32), device=dev) * 0.1
cases.append(("frustum 32->32 k3 (plain)", lambda m: ops.conv(Vol(xv), c32, want_stats=True, math_mode=m)))
cases.append(("frustum 32->32 k3 (pending)", lambda m: ops.conv(Vol(xv, s3, h3, ops.SS_ACT_RELU), c32, want_stats=True, math_mode=m)))
dc = vt.depth_net.depth_conv[0].conv1
xd = torch.randn((1, 1, H, W, 640), device=dev)
cases.append(("depth 640->640 k3 2-D", lambda m: ops.conv(Vol(xd), dc, math_mode=m)))
for name, fn in cases:
    row = []
    for mname, m in (("tf32", ops.SS_MATH_TF32), ("f16", ops.SS_MATH_F16), ("f16x3", ops.SS_MATH_TF32X3)):
        ops.arena(dev).reset()
        row.append(f"{mname} {timeit(lambda: (ops.arena(dev).reset(), fn(m))):.4f} ms")
    print(f"{name:32s} " + "   ".join(row), flush=True)
