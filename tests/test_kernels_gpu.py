"""Per-kernel parity: every C-ABI entry point against the CPU oracle on the same seeded inputs.

Tolerances (north star: <= 1e-3 relative fp32, bit-exact for the voxel-index scatter):
  * SS_MATH_3XTF32 (split TF32, ~fp32):  rel-to-max error <= 2e-4 (tensor-core fp32 accumulation over
    K up to 10368 products is the floor, not the operand split)
  * SS_MATH_TF32   (tensor-core TF32):    rel-to-max error <= 1e-3  (per layer, same inputs)
  * integer / index paths: exact equality
"""
import math
import zlib

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import restatement as O
from util import rel_err

pytestmark = pytest.mark.gpu

TOL = {"precise": 2e-4, "tf32": 1e-3, "tf32x3": 2e-4, "f16": 1e-3}


@pytest.fixture(scope="module")
def ops():
    from stereoscene_b200 import cabi, ops as _ops
    cabi.load()
    return _ops


def _mode(ops, name):
    return {"precise": ops.SS_MATH_3XTF32, "tf32": ops.SS_MATH_TF32, "tf32x3": ops.SS_MATH_TF32X3, "f16": ops.SS_MATH_F16}[name]


def _cl(x):  # NCDHW cpu -> channels-last cuda [B,D,H,W,C]
    return x.permute(0, 2, 3, 4, 1).contiguous().cuda()


def _ncdhw(y):  # channels-last cuda -> NCDHW cpu
    return y.permute(0, 4, 1, 2, 3).cpu()


CONV_CASES = [
    # name, module factory, input shape [B,C,D,H,W]
    ("k3s1_32", lambda: nn.Conv3d(32, 32, 3, 1, 1, bias=False), (2, 32, 6, 9, 11)),
    ("k3s2_32_64", lambda: nn.Conv3d(32, 64, 3, 2, 1, bias=False), (1, 32, 8, 10, 12)),
    ("k3s1_64_128", lambda: nn.Conv3d(64, 128, 3, 1, 1, bias=False), (1, 64, 4, 6, 10)),
    ("k3s1_384_192", lambda: nn.Conv3d(384, 192, 3, 1, 1, bias=False), (1, 384, 4, 8, 6)),
    ("k1_192_20", lambda: nn.Conv3d(192, 20, 1, bias=False), (2, 192, 3, 5, 7)),
    ("k1s2_128_256", lambda: nn.Conv3d(128, 256, 1, 2, bias=False), (1, 128, 4, 6, 8)),
    ("k3_32_1", lambda: nn.Conv3d(32, 1, 3, 1, 1, bias=True), (2, 32, 5, 6, 9)),
    ("k3_2_32_bias", lambda: nn.Conv3d(2, 32, 3, 1, 1, bias=True), (2, 2, 6, 7, 9)),
    ("tk3s2_128_64", lambda: nn.ConvTranspose3d(128, 64, 3, padding=1, output_padding=1, stride=2, bias=False), (1, 128, 3, 4, 5)),
    ("tk3s2_64_32", lambda: nn.ConvTranspose3d(64, 32, 3, padding=1, output_padding=1, stride=2, bias=False), (2, 64, 2, 3, 4)),
    ("tk2s2_256_128", lambda: nn.ConvTranspose3d(256, 128, 2, 2, bias=False), (1, 256, 3, 4, 2)),
    ("tk4s4_512_128", lambda: nn.ConvTranspose3d(512, 128, 4, 4, bias=False), (1, 512, 2, 3, 1)),
    ("tk1s1_128_128", lambda: nn.ConvTranspose3d(128, 128, 1, 1, bias=False), (1, 128, 3, 4, 2)),
    ("k3s1_32_march", lambda: nn.Conv3d(32, 32, 3, 1, 1, bias=True), (2, 32, 11, 21, 19)),
    ("k1_32_32_march", lambda: nn.Conv3d(32, 32, 1, bias=True), (2, 32, 9, 20, 17)),
    ("k3s1_32_march_big", lambda: nn.Conv3d(32, 32, 3, 1, 1, bias=False), (1, 32, 40, 48, 40)),
    ("k3s1_128_128_halo", lambda: nn.Conv3d(128, 128, 3, 1, 1, bias=True), (1, 128, 5, 40, 16)),
    ("k3s1_384_192_halo", lambda: nn.Conv3d(384, 192, 3, 1, 1, bias=False), (1, 384, 3, 32, 8)),
    ("k3s1_64_64_halo", lambda: nn.Conv3d(64, 64, 3, 1, 1, bias=False), (2, 64, 4, 24, 24)),
    ("k3s1_256_256_halo", lambda: nn.Conv3d(256, 256, 3, 1, 1, bias=False), (1, 256, 3, 64, 8)),
    ("tk3s2_64_32_tp", lambda: nn.ConvTranspose3d(64, 32, 3, padding=1, output_padding=1, stride=2, bias=False), (1, 64, 5, 16, 16)),
    ("tk3s2_128_64_tp", lambda: nn.ConvTranspose3d(128, 64, 3, padding=1, output_padding=1, stride=2, bias=True), (2, 128, 3, 12, 24)),
    ("k3s1_256_512", lambda: nn.Conv3d(256, 512, 3, 1, 1, bias=False), (1, 256, 3, 5, 6)),
    ("k3s1_128_128_big", lambda: nn.Conv3d(128, 128, 3, 1, 1, bias=True), (2, 128, 9, 12, 16)),
]


@pytest.mark.parametrize("mode", ["precise", "tf32", "tf32x3", "f16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3d_family(ops, case, mode):
    name, make, shape = case
    torch.manual_seed(zlib.crc32(name.encode()) % 1000)
    m = make()
    x = torch.randn(shape)
    want = m(x).detach()
    mg = make().cuda()
    mg.load_state_dict(m.state_dict())
    y, st = ops.conv(ops.Vol(_cl(x)), mg, want_stats=True, math_mode=_mode(ops, mode))
    torch.cuda.synchronize()
    assert rel_err(_ncdhw(y), want) < TOL[mode]
    # epilogue sums feed GroupNorm: compare with the oracle's sums of the same tensor
    s = want.double().sum(dim=(2, 3, 4))
    q = (want.double() ** 2).sum(dim=(2, 3, 4))
    assert rel_err(st[..., 0], s) < 10 * TOL[mode]
    assert rel_err(st[..., 1], q) < 5 * TOL[mode]


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[0] not in ("k1_192_20", "k3_32_1", "k3_2_32_bias")],
                         ids=[c[0] for c in CONV_CASES if c[0] not in ("k1_192_20", "k3_32_1", "k3_2_32_bias")])
def test_tcgen05_conv_agrees_with_mma_sync(ops, case):
    """The tcgen05 (TMEM / mbarrier pipeline) kernel against the mma.sync kernel, both TF32, with a
    pending affine + ReLU on the input and bias-free epilogue sums."""
    name, make, shape = case
    torch.manual_seed(zlib.crc32(name.encode()) % 1000 + 1)
    mg = make().cuda()
    x = _cl(torch.randn(shape))
    B, Cin = shape[0], shape[1]
    sc = (torch.rand(B, Cin) + 0.5).cuda()
    sh = (torch.randn(B, Cin) * 0.3).cuda()
    v = ops.Vol(x, sc, sh, ops.SS_ACT_RELU)
    try:
        ops.use_tcgen05(False)
        y0, st0 = ops.conv(v, mg, want_stats=True, math_mode=ops.SS_MATH_TF32)
        ops.use_tcgen05(True)
        y1, st1 = ops.conv(v, mg, want_stats=True, math_mode=ops.SS_MATH_TF32)
        torch.cuda.synchronize()
    finally:
        ops.use_tcgen05(True)
    assert rel_err(y1, y0) < 1e-3
    assert rel_err(st1[..., 1], st0[..., 1]) < 2e-3


@pytest.mark.parametrize("single_launch", [True, False], ids=["f16x3", "3launch"])
@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[0] not in ("k3_2_32_bias",)],
                         ids=[c[0] for c in CONV_CASES if c[0] not in ("k3_2_32_bias",)])
def test_compensated_tcgen05_conv_matches_fp32(ops, case, single_launch):
    """SS_MATH_TF32X3 with a pending affine + ReLU on the input, an output activation and the epilogue sums, against an
    fp64 evaluation of the layer.  Two implementations: three accumulating TF32 launches (lo*hi + hi*lo + hi*hi; every
    tcgen05 kernel family) and, on the halo-resident / box kernels, ONE launch with both operands split into fp16 halves
    (SS_MATH_F16X3, six kind::f16 MMAs per chunk)."""
    name, make, shape = case
    ops.use_f16x3(single_launch)
    torch.manual_seed(zlib.crc32(name.encode()) % 1000 + 2)
    m = make()
    x = torch.randn(shape)
    B, Cin = shape[0], shape[1]
    sc, sh = torch.rand(B, Cin) + 0.5, torch.randn(B, Cin) * 0.3
    xin = torch.relu(x * sc.view(B, Cin, 1, 1, 1) + sh.view(B, Cin, 1, 1, 1))
    want = torch.relu(m.double()(xin.double())).float()
    mg = make().cuda()
    mg.load_state_dict(m.float().state_dict())
    v = ops.Vol(_cl(x), sc.cuda(), sh.cuda(), ops.SS_ACT_RELU)
    c0 = ops.cabi.kernel_census()
    try:
        y, st = ops.conv(v, mg, want_stats=True, out_act=ops.SS_ACT_RELU, math_mode=ops.SS_MATH_TF32X3)
        torch.cuda.synchronize()
    finally:
        ops.use_f16x3(True)
    c1 = ops.cabi.kernel_census()
    ran = {k: c1[k] - c0.get(k, 0) for k in c1 if k.startswith("conv_") and c1[k] - c0.get(k, 0) > 0}
    assert "conv_igemm_kernel" not in ran, "compensated mode fell back to the mma.sync kernel"
    f16 = [k for k in ran if "f16x3" in k]
    if single_launch and f16:
        assert sum(ran.values()) == 1, ran                      # ONE launch
    elif not single_launch:
        assert not f16 and sum(ran.values()) == 3, ran          # three accumulating launches
    assert rel_err(_ncdhw(y), want) < 5e-5
    assert rel_err(st[..., 0], want.double().sum(dim=(2, 3, 4))) < 1e-4
    assert rel_err(st[..., 1], (want.double() ** 2).sum(dim=(2, 3, 4))) < 1e-4


def test_single_output_channel_conv(ops):
    """32 -> 1 k3 layers (classif3_2 / MIE redir2) run on the FMA-pipe kernel: pending affine + ReLU in,
    bias + ReLU out, W not a multiple of the 8-voxel segment."""
    torch.manual_seed(21)
    for cin, shape in ((32, (2, 32, 5, 7, 13)), (64, (1, 64, 3, 4, 21))):
        m = nn.Conv3d(cin, 1, 3, 1, 1, bias=True)
        x = torch.randn(shape)
        sc, sh = torch.rand(shape[0], cin) + 0.5, torch.randn(shape[0], cin) * 0.3
        xin = F.relu(x * sc[:, :, None, None, None] + sh[:, :, None, None, None])
        want = F.relu(m(xin)).detach()
        mg = nn.Conv3d(cin, 1, 3, 1, 1, bias=True).cuda()
        mg.load_state_dict(m.state_dict())
        y, _ = ops.conv(ops.Vol(_cl(x), sc.cuda(), sh.cuda(), ops.SS_ACT_RELU), mg, out_act=ops.SS_ACT_RELU,
                        math_mode=ops.SS_MATH_TF32)
        assert rel_err(_ncdhw(y), want) < 1e-3


def test_few_input_channel_conv(ops):
    """Conv3d 2 -> 32 k3 + bias + ReLU (MIE redir1): generic K = tap*Cin path, input as a channel slice."""
    torch.manual_seed(22)
    m = nn.Conv3d(2, 32, 3, 1, 1, bias=True)
    x = torch.rand(2, 2, 6, 9, 13)
    want = F.relu(m(x)).detach()
    mg = nn.Conv3d(2, 32, 3, 1, 1, bias=True).cuda()
    mg.load_state_dict(m.state_dict())
    wide = torch.zeros(2, 6, 9, 13, 4, device="cuda")
    wide[..., 1:3] = _cl(x)                                   # input as a channel slice (ldc = 4)
    y, _ = ops.conv(ops.Vol(wide[..., 1:3]), mg, out_act=ops.SS_ACT_RELU, math_mode=ops.SS_MATH_TF32)
    assert rel_err(_ncdhw(y), want) < 1e-3
    yp, _ = ops.conv(ops.Vol(wide[..., 1:3]), mg, out_act=ops.SS_ACT_RELU, math_mode=ops.SS_MATH_3XTF32)
    assert rel_err(_ncdhw(yp), want) < 2e-4
    # one input channel, 20 output channels (odd row pitch), map wider than one 32-voxel tile, GroupNorm sums
    m1 = nn.Conv3d(1, 20, 3, 1, 1, bias=True)
    x1 = torch.randn(1, 1, 4, 11, 45)
    want1 = F.gelu(m1(x1)).detach()
    mg1 = nn.Conv3d(1, 20, 3, 1, 1, bias=True).cuda()
    mg1.load_state_dict(m1.state_dict())
    ops.arena(torch.device("cuda", 0)).reset()
    y1, st1 = ops.conv(ops.Vol(_cl(x1)), mg1, out_act=ops.SS_ACT_GELU, want_stats=True)
    assert rel_err(_ncdhw(y1), want1) < 1e-3
    wd = want1.double()
    assert rel_err(st1[..., 0], wd.sum(dim=(2, 3, 4))) < 1e-3 and rel_err(st1[..., 1], (wd * wd).sum(dim=(2, 3, 4))) < 1e-3


@pytest.mark.parametrize("mode", ["precise", "tf32"])
def test_conv2d_dilated_bias_gelu(ops, mode):
    torch.manual_seed(3)
    for m in (nn.Conv2d(64, 128, 3, 1, 1), nn.Conv2d(64, 32, 3, padding=6, dilation=6, bias=False), nn.Conv2d(128, 64, 1)):
        x = torch.randn(2, m.in_channels, 12, 20)
        want = F.gelu(m(x)).detach()
        mg = type(m)(m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding, m.dilation, bias=m.bias is not None).cuda()
        mg.load_state_dict(m.state_dict())
        xcl = x.permute(0, 2, 3, 1).contiguous().unsqueeze(1).cuda()          # [B,1,H,W,C]
        y, _ = ops.conv(ops.Vol(xcl), mg, out_act=ops.SS_ACT_GELU, math_mode=_mode(ops, mode))
        got = y.squeeze(1).permute(0, 3, 1, 2).cpu()
        assert rel_err(got, want) < TOL[mode]


@pytest.mark.parametrize("pending", [False, True])
def test_conv2d_3x3_wide_halo_kernel(ops, pending):
    """2-D 3x3 layers with wide channels (DepthNet's BasicBlocks / reduce convs) ride the halo-resident
    tcgen05 kernel with one plane per channel chunk (kd = 1); plain and pending-affine inputs, a map
    whose rows do not fill the 32x8 tile, batch 2."""
    torch.manual_seed(31)
    m = nn.Conv2d(96, 160, 3, 1, 1, bias=True)
    x = torch.randn(2, 96, 30, 19)
    sc, sh = torch.rand(2, 96) + 0.5, torch.randn(2, 96) * 0.3
    xin = F.relu(x * sc[:, :, None, None] + sh[:, :, None, None]) if pending else x
    want = m(xin).detach()
    mg = nn.Conv2d(96, 160, 3, 1, 1, bias=True).cuda()
    mg.load_state_dict(m.state_dict())
    xcl = x.permute(0, 2, 3, 1).contiguous().unsqueeze(1).cuda()
    v = ops.Vol(xcl, sc.cuda(), sh.cuda(), ops.SS_ACT_RELU) if pending else ops.Vol(xcl)
    ops.arena(torch.device("cuda", 0)).reset()
    y, st = ops.conv(v, mg, want_stats=True)
    got = y.squeeze(1).permute(0, 3, 1, 2).cpu()
    assert rel_err(got, want) < TOL["tf32"]
    wd = want.double()
    assert rel_err(st[..., 0], wd.sum(dim=(2, 3))) < 1e-3 and rel_err(st[..., 1], (wd * wd).sum(dim=(2, 3))) < 1e-3


@pytest.mark.parametrize("dims", [(1, 16, 62), (3, 8, 70)])
def test_conv_halo_swapped_axes_and_160_column_tile(ops, dims):
    """Maps that are short in h and long in w run the halo kernel with its 32x8 tile laid along (w,h)
    (tensor-map axes permuted, weight taps transposed); Cout = 160 exercises the 160-column tile.
    2-D (kd = 1) and 3-D, pending affine + ReLU on the input, GroupNorm sums out."""
    D, H, W = dims
    torch.manual_seed(41)
    m = nn.Conv3d(64, 160, (3 if D > 1 else 1, 3, 3), 1, (1 if D > 1 else 0, 1, 1), bias=False)
    x = torch.randn(2, 64, D, H, W)
    sc, sh = torch.rand(2, 64) + 0.5, torch.randn(2, 64) * 0.3
    xin = F.relu(x * sc[:, :, None, None, None] + sh[:, :, None, None, None])
    want = m(xin).detach()
    mg = nn.Conv3d(64, 160, m.kernel_size, 1, m.padding, bias=False).cuda()
    mg.load_state_dict(m.state_dict())
    ops.arena(torch.device("cuda", 0)).reset()
    y, st = ops.conv(ops.Vol(_cl(x), sc.cuda(), sh.cuda(), ops.SS_ACT_RELU), mg, want_stats=True)
    assert rel_err(_ncdhw(y), want) < TOL["tf32"]
    wd = want.double()
    assert rel_err(st[..., 0], wd.sum(dim=(2, 3, 4))) < 1e-3 and rel_err(st[..., 1], (wd * wd).sum(dim=(2, 3, 4))) < 1e-3
    ref, _ = ops.conv(ops.Vol(_cl(x), sc.cuda(), sh.cuda(), ops.SS_ACT_RELU), mg, math_mode=ops.SS_MATH_3XTF32)
    assert rel_err(y, ref) < TOL["tf32"]


@pytest.mark.parametrize("case", ["hg_redir2", "head_classifier", "neck_k1_slice", "input_proj", "neck_k2s2", "neck_k4s4"])
def test_pointwise_streaming_kernel(ops, case):
    """1x1x1 layers with short K on the persistent resident-weight kernel (conv_pw.cu): pending GroupNorm + ReLU
    input, bias, odd Cout (20), ConvTranspose3d k = s = 1 into a channel slice of a wider buffer, GroupNorm sums,
    batch 2 (sums per sample).  Volumes are large enough (>= 2 tiles per SM) to take that kernel."""
    torch.manual_seed(61)
    if case == "hg_redir2":
        m, shape, pending, stats = nn.Conv3d(64, 64, 1, bias=False), (1, 64, 40, 32, 32), True, True
    elif case == "head_classifier":
        m, shape, pending, stats = nn.Conv3d(192, 20, 1, bias=True), (2, 192, 20, 32, 32), True, False
    elif case == "neck_k1_slice":
        m, shape, pending, stats = nn.ConvTranspose3d(128, 128, 1, 1, bias=False), (2, 128, 24, 32, 32), False, True
    elif case == "neck_k2s2":      # 8 output parity classes, weights resident per class
        m, shape, pending, stats = nn.ConvTranspose3d(256, 128, 2, 2, bias=False), (2, 256, 8, 32, 32), False, True
    elif case == "neck_k4s4":      # 64 classes x 2 column halves (K = 512 does not fit with 128 columns)
        m, shape, pending, stats = nn.ConvTranspose3d(512, 128, 4, 4, bias=False), (1, 512, 8, 16, 16), False, True
    else:
        m, shape, pending, stats = nn.Conv3d(128, 128, 1, bias=False), (1, 128, 40, 32, 32), False, True
    B, Cin = shape[0], shape[1]
    x = torch.randn(shape)
    sc, sh = torch.rand(B, Cin) + 0.5, torch.randn(B, Cin) * 0.3
    xin = F.relu(x * sc[:, :, None, None, None] + sh[:, :, None, None, None]) if pending else x
    want = m(xin).detach()
    mg = type(m)(m.in_channels, m.out_channels, m.kernel_size, m.stride, bias=m.bias is not None).cuda()
    mg.load_state_dict(m.state_dict())
    v = ops.Vol(_cl(x), sc.cuda(), sh.cuda(), ops.SS_ACT_RELU) if pending else ops.Vol(_cl(x))
    ops.arena(torch.device("cuda", 0)).reset()
    out = None
    if case.startswith("neck"):
        wide = torch.full((B,) + tuple(want.shape[2:]) + (384,), 7.0, device="cuda")
        out = wide[..., 128:256]
    y, st = ops.conv(v, mg, out=out, want_stats=stats)
    assert rel_err(_ncdhw(y), want) < TOL["tf32"]
    if stats:
        wd = want.double()
        assert rel_err(st[..., 0], wd.sum(dim=(2, 3, 4))) < 1e-3 and rel_err(st[..., 1], (wd * wd).sum(dim=(2, 3, 4))) < 1e-3
    if out is not None:
        assert float(wide[..., :128].min()) == 7.0 and float(wide[..., 256:].max()) == 7.0
    ref, _ = ops.conv(v, mg, math_mode=ops.SS_MATH_3XTF32)
    assert rel_err(y, ref) < TOL["tf32"]


@pytest.mark.parametrize("mode", ["precise", "tf32"])
def test_conv_pending_affine_relu_and_gn_chain(ops, mode):
    """conv -> GroupNorm -> ReLU -> conv -> GroupNorm, with both norms applied as pending affines,
    versus the plain PyTorch composition; also writes into a channel slice of a wider buffer."""
    torch.manual_seed(5)
    c1, g1 = nn.Conv3d(32, 64, 3, 1, 1, bias=False), nn.GroupNorm(2, 64)
    c2, g2 = nn.Conv3d(64, 32, 3, 2, 1, bias=True), nn.GroupNorm(32, 32)
    for g in (g1, g2):
        nn.init.uniform_(g.weight, 0.5, 1.5); nn.init.normal_(g.bias, 0, 0.2)
    x = torch.randn(2, 32, 6, 8, 10)
    want1 = F.relu(g1(c1(x)))
    want2 = g2(c2(want1)).detach()
    mods = [m.cuda() for m in (nn.Conv3d(32, 64, 3, 1, 1, bias=False), nn.GroupNorm(2, 64), nn.Conv3d(64, 32, 3, 2, 1, bias=True), nn.GroupNorm(32, 32))]
    for a, b in zip(mods, (c1, g1, c2, g2)):
        a.load_state_dict(b.state_dict())
    ops.arena(torch.device("cuda", 0)).reset()
    y1, st1 = ops.conv(ops.Vol(_cl(x)), mods[0], want_stats=True, math_mode=_mode(ops, mode))
    v1 = ops.gn_pending(y1, st1, mods[1], ops.SS_ACT_RELU)
    wide = torch.full((2, 3, 4, 5, 48), 7.0, device="cuda")
    y2, st2 = ops.conv(v1, mods[2], out=wide[..., 8:40], want_stats=True, math_mode=_mode(ops, mode))
    v2 = ops.gn_pending(y2, st2, mods[3])
    got = _ncdhw(v2.plain())
    assert rel_err(got, want2) < 3 * TOL[mode]
    assert rel_err(_ncdhw(v1.plain()), want1.detach()) < 3 * TOL[mode]
    assert float(wide[..., :8].min()) == 7.0 and float(wide[..., 40:].max()) == 7.0      # neighbours untouched


@pytest.mark.parametrize("chans", [(64, 32), (128, 64)])
def test_conv_join_fused_in_transposed_epilogue(ops, chans):
    """Hourglass up-convolution + eval-BatchNorm + residual (with its own pending GroupNorm affine) + ReLU as ONE
    kernel (ss_conv3d_tc_join_fwd) against the PyTorch composition, and against the unfused conv + join path."""
    from stereoscene_b200 import cabi
    cin, cout = chans
    torch.manual_seed(51)
    m = nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False)
    bn = nn.BatchNorm3d(cout).eval()
    nn.init.uniform_(bn.weight, 0.5, 1.5); nn.init.normal_(bn.bias, 0, 0.2)
    bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
    x = torch.randn(2, cin, 3, 16, 15)
    sc, sh = torch.rand(2, cin) + 0.5, torch.randn(2, cin) * 0.3
    r = torch.randn(2, cout, 6, 32, 30)
    rs, rh = torch.rand(2, cout) + 0.5, torch.randn(2, cout) * 0.3
    xin = F.relu(x * sc[:, :, None, None, None] + sh[:, :, None, None, None])
    want = F.relu(bn(m(xin)) + (r * rs[:, :, None, None, None] + rh[:, :, None, None, None])).detach()
    mg = nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False).cuda()
    mg.load_state_dict(m.state_dict())
    want64 = F.relu(bn.double()(m.double()(xin.double())) + (r * rs[:, :, None, None, None] + rh[:, :, None, None, None]).double()).float().detach()
    bn.float(); m.float()
    xv = ops.Vol(_cl(x), sc.cuda(), sh.cuda(), ops.SS_ACT_RELU)
    rv = ops.Vol(_cl(r), rs.cuda(), rh.cuda(), ops.SS_ACT_NONE)
    d = cabi.ConvDesc(2, 3, 16, 15, cin, 6, 32, 30, cout, 3, 3, 3, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, cin, cout, 1, 1, 0, max(32, cout))
    assert cabi.load().ss_conv3d_tc_join_supported(d) == 1                      # this shape takes the fused kernel
    got = ops.conv_join(xv, mg, ops.bn_pending(_cl(x), bn.cuda()), rv, out_act=ops.SS_ACT_RELU)
    assert rel_err(_ncdhw(got), want) < TOL["tf32"]
    y, _ = ops.conv(xv, mg)
    unfused = ops.join(ops.bn_pending(y, bn), rv, out_act=ops.SS_ACT_RELU)
    assert rel_err(got, unfused) < 1e-5
    # the compensated mode runs the same fused kernel three times (the join belongs to the last pass)
    ops.set_default_math(ops.SS_MATH_TF32X3)
    try:
        for single, kern, launches in ((False, "conv_tpose_kernel", 3), (True, "conv_tpose_f16x3_kernel", 1)):
            ops.use_f16x3(single)
            n0 = cabi.kernel_census().get(kern, 0)
            got3 = ops.conv_join(xv, mg, ops.bn_pending(_cl(x), bn.cuda()), rv, out_act=ops.SS_ACT_RELU)
            torch.cuda.synchronize()
            assert cabi.kernel_census()[kern] == n0 + launches
            assert rel_err(_ncdhw(got3), want64) < 5e-5, single
    finally:
        ops.use_f16x3(True)
        ops.set_default_math(ops.SS_MATH_TF32)
    # a layer the fused kernel does not take falls back to conv + join
    m2 = nn.ConvTranspose3d(64, 32, 2, 2, bias=False).cuda()
    x2 = torch.randn(1, 64, 2, 4, 4)
    got2 = ops.conv_join(ops.Vol(_cl(x2)), m2, None, None, out_act=ops.SS_ACT_RELU)
    assert rel_err(_ncdhw(got2), F.relu(m2.cpu()(x2)).detach()) < TOL["tf32"]


def test_bn_pending_join_and_alpha(ops):
    torch.manual_seed(6)
    bn = nn.BatchNorm3d(16).eval()
    nn.init.uniform_(bn.weight, 0.5, 1.5); nn.init.normal_(bn.bias, 0, 0.2)
    bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
    x, r = torch.randn(2, 16, 3, 5, 7), torch.randn(2, 16, 3, 5, 7)
    alpha = torch.tensor([0.37])
    want = F.relu(alpha * bn(x) + F.relu(r)).detach()
    got = ops.join(ops.bn_pending(_cl(x), bn.cuda()), ops.Vol(_cl(r), None, None, ops.SS_ACT_RELU),
                   out_act=ops.SS_ACT_RELU, alpha=alpha.cuda())
    assert rel_err(_ncdhw(got), want) < 1e-6
    # scalar (C not a multiple of 4) path
    x3, r3 = torch.randn(1, 3, 2, 3, 5), torch.randn(1, 3, 2, 3, 5)
    got3 = ops.join(ops.Vol(_cl(x3)), ops.Vol(_cl(r3)))
    assert rel_err(_ncdhw(got3), x3 + r3) < 1e-6


def test_softmax_over_depth(ops):
    torch.manual_seed(7)
    x = torch.randn(2, 240, 6, 10) * 3
    got = ops.softmax_d(x.cuda()[:, :112])             # channel slice of a wider NCHW tensor
    assert rel_err(got, F.softmax(x[:, :112], dim=1)) < 1e-6
    assert rel_err(got.sum(1), torch.ones(2, 6, 10)) < 1e-6


def test_ca3d_block(ops):
    """CA3D (attention.py:113-120) with the gate folded into the pending affine."""
    from stereoscene_b200.plugin.view_transformer import CA3DParams
    from stereoscene_b200 import synth
    torch.manual_seed(8)
    fn = CA3DParams(32)
    synth.randomize_weights_(fn, 3)
    sd = {"p." + k: v for k, v in fn.state_dict().items()}
    x = torch.randn(2, 32, 6, 8, 10)
    want = O.ca3d(sd, "p", x)
    fn = fn.cuda()
    ops.arena(torch.device("cuda", 0)).reset()
    xv = ops.Vol(_cl(x))
    d, st = ops.conv(xv, fn.conv1[0], out_act=ops.SS_ACT_GELU, want_stats=True, math_mode=ops.SS_MATH_3XTF32)
    dv = ops.ca3d_gate(ops.gn_pending(d, st, fn.conv1[2]), st, fn.conv2[0], fn.conv2[2])
    o, st2 = ops.conv(dv, fn.conv[0], out_act=ops.SS_ACT_GELU, want_stats=True, math_mode=ops.SS_MATH_3XTF32)
    got = ops.gn_pending(o, st2, fn.conv[2]).plain()
    assert rel_err(_ncdhw(got), want) < 2e-4


@pytest.mark.parametrize("shape", [(2, 64, 8, 16, 48), (1, 64, 6, 40, 112), (1, 32, 4, 12, 20)])
def test_gwc_warp(ops, shape):
    """Fused correlation + warp vs build_gwc_volume + warp of the oracle; includes bins whose
    disparity falls beyond the image / beyond maxdisp (zero) and per-sample calibration."""
    B, C, H, W, D = shape
    torch.manual_seed(9)
    ref, tgt = torch.randn(B, C, H, W), torch.randn(B, C, H, W)
    calib = torch.tensor([[380.3], [141.0]])[:B]
    want = O.warp_disparity_to_depth(O.gwc_volume(ref, tgt, D, 32), calib)        # [B,32,D,H,W]
    fea = torch.cat([ref, tgt], 0).permute(0, 2, 3, 1).contiguous().unsqueeze(1).cuda()
    got = ops.gwc_warp(fea, calib.cuda(), D, 32)                                   # [B,D,H,W,32]
    assert rel_err(got.permute(0, 4, 1, 2, 3), want) < 1e-5


@pytest.mark.parametrize("mode", ["precise", "tf32"])
@pytest.mark.parametrize("dims", [(2, 48, 8, 16), (1, 112, 6, 23), (1, 20, 5, 13), (1, 48, 16, 48), (2, 112, 24, 80),
                                  (1, 20, 8, 20), (2, 100, 10, 18), (1, 112, 48, 160)])
def test_bri_attention(ops, mode, dims):
    B, D, H, W = dims
    torch.manual_seed(10)
    q = F.softmax(torch.randn(B, D, H, W) * 2, dim=1)
    kv = F.softmax(torch.randn(B, D, H, W) * 2, dim=1)
    sd = {"a.query_conv.weight": torch.tensor(6.5).view(1, 1, 1, 1, 1), "a.query_conv.bias": torch.tensor([0.1]),
          "a.key_conv.weight": torch.tensor(5.5).view(1, 1, 1, 1, 1), "a.key_conv.bias": torch.tensor([-0.05]),
          "a.value_conv.weight": torch.tensor(1.3).view(1, 1, 1, 1, 1), "a.value_conv.bias": torch.tensor([0.02]),
          "a.gamma": torch.tensor([0.5])}
    want = O.bri_attention(sd, "a", q.unsqueeze(1), kv.unsqueeze(1)).squeeze(1)
    params = torch.tensor([6.5, 0.1, 5.5, -0.05, 1.3, 0.02, 0.5]).cuda()
    both = torch.zeros(B, D, H, W, 2, device="cuda")
    ops.bri_attention(q.cuda(), kv.cuda(), params, both[..., 1], 2, math_mode=_mode(ops, mode))
    assert rel_err(both[..., 1], want) < (1e-4 if mode == "precise" else 1e-3)
    assert float(both[..., 0].abs().max()) == 0.0


def _frustum_case(B=2, input_size=(64, 128), dbound=(2.0, 26.0, 0.5), grid=((0.0, 51.2, 3.2), (-25.6, 25.6, 3.2), (-2.0, 4.4, 1.6))):
    from stereoscene_b200 import synth
    left, _, _ = synth.kitti_calibration(B, input_size)
    fr = O.create_frustum(input_size, 8, list(dbound))
    geom = O.get_geometry(fr, left["rots"], left["trans"], left["intrins"], left["post_rots"], left["post_trans"], left["bda"])
    dx, bx, nx = O.gen_dx_bx(*[list(g) for g in grid])
    return geom, dx, bx, nx


def test_voxel_index_bit_exact(ops):
    geom, dx, bx, nx = _frustum_case()
    # add adversarial points: exactly on voxel faces, slightly negative, far outside, huge
    extra = torch.tensor([[0.0, -25.6, -2.0], [-1e-3, 0.0, 0.0], [-3.19, 0.1, 0.1], [51.2, 25.6, 4.4],
                          [3.2, -22.4, -0.4], [1e9, -1e9, 3.0], [51.199997, 25.599998, 4.3999996]])
    g = geom.reshape(2, -1, 3)
    g[0, :extra.shape[0]] = extra
    want_idx, want_kept = O.voxel_indices(g, dx, bx, nx)
    idx = ops.splat_build_index(g.cuda(), dx.tolist(), bx.tolist(), nx.tolist(), want_coords=True)
    c = idx.coords.cpu()
    assert torch.equal(c[:, 3].bool(), want_kept)
    inside = want_kept
    assert torch.equal(c[inside, :3].long(), want_idx[inside])
    # CSR structure: counts per voxel equal the oracle's histogram; order is stable in point id
    n = [int(v) for v in nx.tolist()]
    P = g.shape[1]
    b_ix = torch.arange(2).repeat_interleave(P)
    rank = ((b_ix * n[0] + want_idx[:, 0]) * n[1] + want_idx[:, 1]) * n[2] + want_idx[:, 2]
    hist = torch.bincount(rank[inside], minlength=2 * n[0] * n[1] * n[2])
    start = idx.voxel_start.cpu().long()
    assert torch.equal(start[1:] - start[:-1], hist)
    order = idx.order.cpu().long()[: int(inside.sum())]
    assert torch.equal(rank[order], torch.sort(rank[inside], stable=True)[0])
    seg_sorted = torch.sort(rank[inside], stable=True)[1]
    assert torch.equal(order, torch.nonzero(inside).flatten()[seg_sorted])


def test_lift_splat_matches_oracle(ops):
    geom, dx, bx, nx = _frustum_case()
    B, _, D, H, W, _ = geom.shape
    torch.manual_seed(11)
    dp = F.softmax(torch.randn(B, D, H, W), dim=1)
    feat = torch.randn(B, 128, H, W)
    want = O.lift_splat(dp, feat, geom, dx, bx, nx)                                  # [B,C,X,Y,Z]
    idx = ops.splat_build_index(geom.reshape(B, -1, 3).cuda(), dx.tolist(), bx.tolist(), nx.tolist())
    got = ops.lift_splat(dp.cuda(), feat.permute(0, 2, 3, 1).contiguous().cuda(), idx)    # [B,X,Y,Z,C]
    got = got.permute(0, 4, 1, 2, 3).cpu()
    assert torch.equal(got != 0, want != 0)                    # occupancy mask: exact
    assert rel_err(got, want) < 1e-6
    # checksum of checksums: total mass per channel is conserved by the scatter
    _, kept = O.voxel_indices(geom, dx, bx, nx)
    lifted = (dp.unsqueeze(1) * feat.unsqueeze(2)).permute(0, 2, 3, 4, 1).reshape(-1, 128)[kept]
    assert rel_err(got.double().sum(dim=(0, 2, 3, 4)), lifted.double().sum(0)) < 1e-5


def test_bev_pool_dropin(ops):
    """Same call as the reference's bev_pool(x, geom_feats, B, nz, nx, ny) incl. 0-d tensor sizes,
    ragged occupancy, collisions and an empty point set."""
    torch.manual_seed(12)
    B, Dz, Hx, Wy, C, N = 2, 4, 16, 12, 24, 5000
    coords = torch.stack([torch.randint(0, Hx, (N,)), torch.randint(0, Wy, (N,)), torch.randint(0, Dz, (N,)),
                          torch.randint(0, B, (N,))], 1)
    coords[:300] = coords[0]                                   # heavy collision in one voxel
    feats = torch.randn(N, C)
    want = O.bev_pool(feats, coords, B, Dz, Hx, Wy)
    got = ops.bev_pool(feats.cuda(), coords.cuda(), torch.tensor(float(B)), torch.tensor(float(Dz)),
                       torch.tensor(float(Hx)), torch.tensor(float(Wy)))
    assert got.shape == want.shape and got.is_contiguous()
    assert rel_err(got, want) < 1e-6
    empty = ops.bev_pool(feats[:0].cuda(), coords[:0].cuda(), B, Dz, Hx, Wy)
    assert float(empty.abs().sum()) == 0.0


@pytest.mark.parametrize("mode", ["precise", "tf32"])
def test_deformable_conv(ops, mode):
    """DCN (mmcv DeformConv2dPack semantics) = ss_deform_sample_fwd + grouped GEMM, against
    torchvision's deform_conv2d on the CPU with non-trivial offsets (incl. samples outside the map)."""
    from torchvision.ops import deform_conv2d
    from stereoscene_b200.plugin.view_transformer import DCN
    torch.manual_seed(14)
    m = DCN(128, 128, 3, 1, 4)
    nn.init.normal_(m.conv_offset.weight, 0, 0.05)
    nn.init.normal_(m.conv_offset.bias, 0, 1.5)
    x = torch.randn(2, 128, 9, 14)
    with torch.no_grad():
        want = deform_conv2d(x, m.conv_offset(x), m.weight, None, 1, 1, 1)
    mg = m.cuda()
    ops.set_default_math(_mode(ops, mode))
    try:
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            got = mg(x.cuda())
    finally:
        ops.set_default_math(ops.SS_MATH_TF32)
    assert got.shape == want.shape
    # tf32: the offsets themselves are a TF32 conv output (as in the reference's own GPU run, cuDNN allow_tf32), and
    # this case exaggerates them (|offset| ~ 2.5 px): a ~1e-3 px sampling error times the feature gradient adds to
    # the GEMM's own rounding
    assert rel_err(got, want) < (TOL[mode] if mode == "precise" else 3e-3)


def test_channel_sums(ops):
    """ss_channel_sums_fwd: per-(batch,channel) sum / sum of squares of a pending volume, also on a
    channel slice of a wider buffer."""
    torch.manual_seed(21)
    for (B, V, Cc, ld) in ((2, 8 * 16, 640, 640), (1, 77, 33, 40), (3, 5, 4, 4)):
        wide = torch.randn(B, 1, 1, V, ld)
        x = wide[..., :Cc]
        sc, sh = torch.rand(B, Cc) + 0.5, torch.randn(B, Cc) * 0.3
        a = F.relu(x * sc.view(B, 1, 1, 1, Cc) + sh.view(B, 1, 1, 1, Cc)).double()
        want = torch.stack((a.sum(dim=(1, 2, 3)), (a * a).sum(dim=(1, 2, 3))), -1)
        ops.arena(torch.device("cuda", 0)).reset()
        got = ops.channel_sums(ops.Vol(wide.cuda()[..., :Cc], sc.cuda(), sh.cuda(), ops.SS_ACT_RELU))
        assert got.dtype == torch.float64 and rel_err(got, want) < 1e-6
        plain = ops.channel_sums(ops.Vol(wide.cuda()[..., :Cc]))
        assert rel_err(plain[..., 0], x.double().sum(dim=(1, 2, 3))) < 1e-5


@pytest.mark.parametrize("mode", ["precise", "tf32"])
def test_depth_net_module(ops, mode):
    """DepthNet with the reference's tensor contract: 3 BasicBlocks with pending
    BatchNorm, ASPP (dilations 6/12/18 larger than the map, pooled branch folded into a shift),
    DCN and the 1x1 heads, against the oracle restatement (pinned to the reference's forward)."""
    from stereoscene_b200 import presets, synth
    model, mc = presets.build("tiny")
    synth.randomize_weights_(model, 11)
    vt = model.img_view_transformer
    sd = {"dn." + k: v.detach().cpu().clone() for k, v in vt.depth_net.state_dict().items()}
    B, H, W = 2, 6, 20
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, vt.numC_input, H, W, generator=g)
    mlp = torch.randn(B, 1, vt.cam_channels, generator=g)
    with torch.no_grad():
        want = O.depth_net(sd, "dn", x, mlp)
    dn = vt.depth_net.cuda()
    ops.set_default_math(_mode(ops, mode))
    try:
        ops.arena(torch.device("cuda", 0)).reset()
        with torch.no_grad():
            got = dn(x.cuda(), mlp.cuda())
    finally:
        ops.set_default_math(ops.SS_MATH_TF32)
    assert got.shape == want.shape
    D = vt.D
    tol = 1e-3 if mode == "precise" else 1e-2          # ~12 stacked layers in TF32
    assert rel_err(got[:, :D], want[:, :D]) < tol
    assert rel_err(got[:, D:], want[:, D:]) < tol


def test_trilinear_upsample_and_argmax(ops):
    torch.manual_seed(13)
    for shape, size in (((2, 20, 8, 8, 4), (16, 16, 8)), ((1, 20, 5, 6, 3), (10, 12, 6)), ((1, 7, 4, 4, 2), (9, 7, 5))):
        x = torch.randn(shape)
        want = F.interpolate(x, size=size, mode="trilinear", align_corners=False)
        got, labels = ops.trilinear(_cl(x), size, want_labels=True)
        assert rel_err(_ncdhw(got), want) < 1e-6
        assert torch.equal(labels.cpu().long(), got.argmax(dim=-1).cpu())


def test_layout_round_trip(ops):
    x = torch.randn(2, 37, 5, 9).cuda()
    y = ops.to_channels_last(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.to_channels_first(y), x)


def test_ssc_metrics_kernel_matches_reference(ops):
    """ss_ssc_confusion_fwd through the SSCMetrics module (reference class name / methods): exact
    integer agreement with the reference's own results (golden) and with the oracle on a full-size
    256x256x32 label volume; uint8 and int64 targets, masks, accumulation over samples."""
    import os
    from stereoscene_b200.plugin.metrics import SSCMetrics
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ssc.npz"))
    metric = SSCMetrics().cuda()
    for name in ("c0", "c1", "c2"):
        pred, true = torch.from_numpy(g[name + "_pred"]).cuda(), torch.from_numpy(g[name + "_true"]).cuda()
        ne = torch.from_numpy(g[name + "_nonempty"]).cuda() if name + "_nonempty" in g else None
        ns = torch.from_numpy(g[name + "_nonsurface"]).cuda() if name + "_nonsurface" in g else None
        for tgt in (true, true.long()):                                         # uint8 and int64 ground truth
            tp, fp, fn, tps, fps, fns = SSCMetrics().cuda().compute_single(pred.long(), tgt, ne, ns)
            assert np.array_equal(np.array([tp, fp, fn], dtype=np.float64), g[name + "_completion"])
            assert np.array_equal(tps, g[name + "_tps"]) and np.array_equal(fps, g[name + "_fps"]) and np.array_equal(fns, g[name + "_fns"])
        metric.update(pred, true, ne, ns)
    res = metric.compute()
    assert np.allclose(res["iou_ssc"].cpu().numpy(), g["all_iou_ssc"], rtol=1e-6)
    assert np.allclose([float(res["precision"]), float(res["recall"]), res["iou"], res["iou_ssc_mean"]], g["all_scalars"], rtol=1e-6)
    assert set(metric.state_dict()) == {"tps", "fps", "fns", "completion_tp", "completion_fp", "completion_fn"}
    # full-size volume against the oracle
    gen = torch.Generator().manual_seed(3)
    shape = (1, 256, 256, 32)
    pred = torch.randint(0, 20, shape, generator=gen, dtype=torch.uint8)
    true = torch.randint(0, 20, shape, generator=gen, dtype=torch.uint8)
    true[torch.rand(shape, generator=gen) < 0.1] = 255
    pred[torch.rand(shape, generator=gen) < 0.5] = 0
    want = O.ssc_scores(pred, true)
    comp, tps, fps, fns = SSCMetrics().cuda().scores(pred.cuda(), true.cuda())
    assert torch.equal(comp.cpu(), want["completion"]) and torch.equal(tps.cpu(), want["tps"])
    assert torch.equal(fps.cpu(), want["fps"]) and torch.equal(fns.cpu(), want["fns"])
    # empty input is a no-op
    z = ops.ssc_confusion(pred.cuda()[:0], true.cuda()[:0], 20)
    assert int(z.sum()) == 0
    # predictions outside the class range: misses of their target class, false positives of no class (the oracle's loops agree)
    pred2 = pred.clone()
    pred2[torch.rand(shape, generator=gen) < 0.05] = 37
    want = O.ssc_scores(pred2, true)
    comp, tps, fps, fns = SSCMetrics().cuda().scores(pred2.cuda(), true.cuda())
    assert torch.equal(comp.cpu(), want["completion"]) and torch.equal(tps.cpu(), want["tps"])
    assert torch.equal(fps.cpu(), want["fps"]) and torch.equal(fns.cpu(), want["fns"])


def test_errors_are_loud(ops):
    with pytest.raises(RuntimeError):
        ops.conv(ops.Vol(torch.randn(1, 2, 2, 2, 32)), nn.Conv3d(32, 32, 3, 1, 1))           # CPU tensor
    with pytest.raises(RuntimeError):
        ops.softmax_d(torch.randn(1, 4, 3, 3, dtype=torch.float64).cuda())                     # wrong dtype
    with pytest.raises(RuntimeError):
        ops.conv(ops.Vol(torch.randn(1, 2, 2, 2, 16).cuda()), nn.Conv3d(32, 32, 3, 1, 1).cuda())  # channel mismatch
