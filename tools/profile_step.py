#!/usr/bin/env python
"""Profiling harness for ncu (B200_PROFILING.md): runs warm-up steps, then ONE step (or one named
kernel) between cudaProfilerStart/Stop so `ncu --profile-from-start off` sees exactly that region.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --what step
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/head \
      python tools/profile_step.py --what head_conv
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from stereoscene_b200 import ops, presets, synth  # noqa: E402
from stereoscene_b200.ops import Vol  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="step", choices=["step", "head_conv", "frustum_conv", "enc_conv", "bri", "gwc", "splat", "redir1x1", "depth_conv", "aspp_dil", "mie_redir1", "frustum_conv_pending", "stage2_conv", "tpose32", "hg_conv1", "neck_k4", "hg_redir2", "stage2_conv_plain", "hg_conv4", "hg_conv4_plain", "hg_conv3", "hg_conv3_plain", "hourglass", "pw_proj"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--math", default="mixed", help="math policy (ops.MATH_POLICIES); single-kernel cases run plain TF32 unless --kernel-math says otherwise")
    ap.add_argument("--kernel-math", default="tf32", choices=["tf32", "f16", "tf32x3", "3xtf32"])
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--time-one", action="store_true")
    ap.add_argument("--time", action="store_true", help="time every kernel case with CUDA events instead of profiling one")
    a = ap.parse_args()
    if a.time:
        import subprocess
        for w in ("head_conv", "enc_conv", "frustum_conv", "frustum_conv_pending", "redir1x1", "mie_redir1", "depth_conv", "aspp_dil", "bri", "gwc", "splat"):
            subprocess.run([sys.executable, __file__, "--what", w, "--time-one", "--workload", a.workload])
        return
    dev = torch.device("cuda", 0)
    if a.what == "step":
        ops.set_math_policy(a.math)
    else:
        ops.set_default_math(ops.MATH_MODES[a.kernel_math])
    model, mc = presets.build(a.workload)
    synth.randomize_weights_(model, 0)
    model = model.to(dev).eval()
    xl, xr = synth.stereo_features(1, mc["input_size"], 8, seed=0, device=dev)
    left, right, calib = synth.kitti_calibration(1, mc["input_size"], device=dev)
    vt = model.img_view_transformer
    nx = [int(round(float(v))) for v in vt.nx.detach().cpu()]
    D, H, W = vt.D, vt.frustum.shape[1], vt.frustum.shape[2]

    if a.what == "step":
        fn = lambda: model.forward_features(xl, xr, left, right, calib, occ_size=mc["occ_size"], want_labels=True)   # noqa: E731
    elif a.what == "head_conv":
        head = model.pts_bbox_head.occ_convs[0][0]
        x = torch.randn((1, nx[0], nx[1], nx[2], 384), device=dev)
        sc, sh = torch.rand((1, 384), device=dev) + 0.5, torch.randn((1, 384), device=dev) * 0.1
        y = torch.empty((1, nx[0], nx[1], nx[2], 192), device=dev)
        fn = lambda: ops.conv(Vol(x, sc, sh, ops.SS_ACT_RELU), head, out=y, want_stats=True)    # noqa: E731
    elif a.what == "enc_conv":
        c = model.img_bev_encoder_backbone.layers[0][0].conv1
        x = torch.randn((1, nx[0], nx[1], nx[2], 128), device=dev)
        fn = lambda: ops.conv(Vol(x), c, want_stats=True)     # noqa: E731
    elif a.what == "frustum_conv":
        c = vt.stereo_volume_net.dres0[0][0]
        x = torch.randn((1, D, H, W, 32), device=dev)
        fn = lambda: ops.conv(Vol(x), c, want_stats=True)     # noqa: E731
    elif a.what == "frustum_conv_pending":
        c = vt.stereo_volume_net.dres0[0][0]
        x = torch.randn((1, D, H, W, 32), device=dev)
        sc, sh = torch.rand((1, 32), device=dev) + 0.5, torch.randn((1, 32), device=dev) * 0.1
        fn = lambda: ops.conv(Vol(x, sc, sh, ops.SS_ACT_RELU), c, want_stats=True)     # noqa: E731
    elif a.what == "stage2_conv":
        c = model.img_bev_encoder_backbone.layers[2][1].conv1
        x = torch.randn((1, nx[0] // 4, nx[1] // 4, nx[2] // 4, 512), device=dev)
        sc, sh = torch.rand((1, 512), device=dev) + 0.5, torch.randn((1, 512), device=dev) * 0.1
        fn = lambda: ops.conv(Vol(x, sc, sh, ops.SS_ACT_RELU), c, want_stats=True)     # noqa: E731
    elif a.what == "stage2_conv_plain":
        c = model.img_bev_encoder_backbone.layers[2][1].conv1
        x = torch.randn((1, nx[0] // 4, nx[1] // 4, nx[2] // 4, 512), device=dev)
        fn = lambda: ops.conv(Vol(x), c, want_stats=True)     # noqa: E731
    elif a.what in ("hg_conv4", "hg_conv4_plain"):
        c = vt.stereo_volume_net.dres2.conv4[0][0]
        x = torch.randn((1, D // 4, H // 4, W // 4, 128), device=dev)
        sc, sh = torch.rand((1, 128), device=dev) + 0.5, torch.randn((1, 128), device=dev) * 0.1
        v = Vol(x) if a.what.endswith("plain") else Vol(x, sc, sh, ops.SS_ACT_RELU)
        fn = lambda: ops.conv(v, c, want_stats=True)     # noqa: E731
    elif a.what in ("hg_conv3", "hg_conv3_plain"):
        c = vt.stereo_volume_net.dres2.conv3[0][0]
        x = torch.randn((1, D // 2, H // 2, W // 2, 64), device=dev)
        sc, sh = torch.rand((1, 64), device=dev) + 0.5, torch.randn((1, 64), device=dev) * 0.1
        v = Vol(x) if a.what.endswith("plain") else Vol(x, sc, sh, ops.SS_ACT_RELU)
        fn = lambda: ops.conv(v, c, want_stats=True)     # noqa: E731
    elif a.what == "hourglass":
        from stereoscene_b200.plugin.layers import hourglass
        x = torch.randn((1, D, H, W, 32), device=dev)
        fn = lambda: hourglass(vt.stereo_volume_net.dres2, Vol(x))     # noqa: E731
    elif a.what == "pw_proj":
        c = model.img_bev_encoder_backbone.input_proj[0]
        x = torch.randn((1, nx[0], nx[1], nx[2], 128), device=dev)
        fn = lambda: ops.conv(Vol(x), c, want_stats=True)     # noqa: E731
    elif a.what == "neck_k4":
        c = model.img_bev_encoder_neck.deblocks[2][0]
        x = torch.randn((1, nx[0] // 4, nx[1] // 4, nx[2] // 4, 512), device=dev)
        fn = lambda: ops.conv(Vol(x), c, want_stats=True)     # noqa: E731
    elif a.what == "tpose32":
        c = vt.stereo_volume_net.dres2.conv6[0]
        x = torch.randn((1, D // 2, H // 2, W // 2, 64), device=dev)
        fn = lambda: ops.conv(Vol(x), c)     # noqa: E731
    elif a.what == "hg_conv1":
        c = vt.stereo_volume_net.dres2.conv1[0][0]
        x = torch.randn((1, D, H, W, 32), device=dev)
        fn = lambda: ops.conv(Vol(x), c, want_stats=True)     # noqa: E731
    elif a.what == "hg_redir2":
        c = vt.stereo_volume_net.dres2.redir2[0]
        x = torch.randn((1, D // 2, H // 2, W // 2, 64), device=dev)
        sc, sh = torch.rand((1, 64), device=dev) + 0.5, torch.randn((1, 64), device=dev) * 0.1
        fn = lambda: ops.conv(Vol(x, sc, sh, ops.SS_ACT_RELU), c, want_stats=True)     # noqa: E731
    elif a.what == "redir1x1":
        c = vt.stereo_volume_net.dres2.redir1[0]
        x = torch.randn((1, D, H, W, 32), device=dev)
        fn = lambda: ops.conv(Vol(x), c, want_stats=True)     # noqa: E731
    elif a.what == "mie_redir1":
        c = vt.volume_interaction.redir1
        x = torch.randn((1, D, H, W, 2), device=dev)
        fn = lambda: ops.conv(Vol(x), c, out_act=ops.SS_ACT_RELU)     # noqa: E731
    elif a.what == "depth_conv":
        c = vt.depth_net.depth_conv[0].conv1
        x = torch.randn((1, 1, H, W, 640), device=dev)
        fn = lambda: ops.conv(Vol(x), c)     # noqa: E731
    elif a.what == "aspp_dil":
        c = vt.depth_net.depth_conv[3].aspp3.atrous_conv
        x = torch.randn((1, 1, H, W, 640), device=dev)
        fn = lambda: ops.conv(Vol(x), c)     # noqa: E731
    elif a.what == "bri":
        q = torch.softmax(torch.randn((1, D, H, W), device=dev), 1)
        kv = torch.softmax(torch.randn((1, D, H, W), device=dev), 1)
        out = torch.empty((1, D, H, W, 2), device=dev)
        p = vt.volume_interaction.lss2stereo.packed()
        fn = lambda: ops.bri_attention(q, kv, p, out[..., 0], 2)     # noqa: E731
    elif a.what == "gwc":
        fea = torch.randn((2, 1, H, W, 64), device=dev)
        fn = lambda: ops.gwc_warp(fea, calib, D, 32)     # noqa: E731
    else:
        idx = vt.splat_index(*[left[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")])
        dp = torch.softmax(torch.randn((1, D, H, W), device=dev), 1)
        ft = torch.randn((1, H, W, 128), device=dev)
        fn = lambda: ops.lift_splat(dp, ft, idx)     # noqa: E731

    if a.time_one:
        with torch.no_grad():
            for _ in range(3):
                ops.arena(dev).reset(); fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.arena(dev).reset(); fn()
            e1.record(); torch.cuda.synchronize()
        print(f"[time] {a.what:22s} {e0.elapsed_time(e1) / 20 * 1e3:9.1f} us", flush=True)
        return
    with torch.no_grad():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(a.reps):
            fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profiled", a.what)


if __name__ == "__main__":
    main()
