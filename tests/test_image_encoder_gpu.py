"""Row N2 on the GPU: the image encoder's kernels through the C-ABI against the CPU oracle (oracle/restatement_image.py)
and against the fixtures the reference's own efficientnet.py wrote.

Tolerances: the CUDA-core kernels (stem, depthwise, SE) are fp32 -> 1e-5 rel-to-max; the whole encoder under the product
math policy (group "image" compensated) stays within 1e-4 of the reference in both error norms (the north star's bar is
1e-3 on the final logits; the encoder sits in front of the volumetric path and must not eat that budget); plain TF32 is
held to 5e-3.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import restatement_image as RI
from util import build_image_encoder, cpu_state_dict, golden_image, image_inputs, rel_err, stage_error

pytestmark = pytest.mark.gpu

LEVELS = ("img_level2", "img_level3", "img_level4", "img_level5", "img_level6", "img_feat")


@pytest.fixture(scope="module")
def ops():
    from stereoscene_b200 import cabi, ops as _ops
    cabi.load()
    return _ops


def _nhwc(x):      # NCHW cpu -> [N,1,H,W,C] cuda
    return x.permute(0, 2, 3, 1).contiguous().unsqueeze(1).cuda()


def _nchw(y):      # [N,1,H,W,C] cuda -> NCHW cpu
    return y.squeeze(1).permute(0, 3, 1, 2).cpu()


@pytest.mark.parametrize("size", [(64, 128), (37, 51), (9, 16)])
def test_stem_conv_matches_oracle(ops, size):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, *size, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(64, generator=g) * 0.1
    want = RI.swish(RI.conv_same(x, w, 2) + b.view(1, -1, 1, 1))
    wp = w.permute(2, 3, 1, 0).reshape(27, 64).contiguous().cuda()
    got = ops.stem_conv2d(x.cuda(), wp, b.cuda(), 3, 2, ops.SS_ACT_SWISH)
    assert rel_err(_nchw(got), want) < 1e-5


@pytest.mark.parametrize("k,s,shape", [(3, 1, (2, 64, 24, 40)), (3, 2, (1, 192, 20, 36)), (5, 1, (2, 288, 13, 17)),
                                       (5, 2, (1, 480, 12, 40)), (5, 2, (2, 96, 11, 9)), (3, 1, (1, 3840, 12, 40)),
                                       (3, 2, (1, 32, 5, 3))])
def test_depthwise_conv_and_pool_match_oracle(ops, k, s, shape):
    g = torch.Generator().manual_seed(k * 10 + s)
    x = torch.randn(*shape, generator=g)
    Cc = shape[1]
    w = torch.randn(Cc, 1, k, k, generator=g) * 0.3
    b = torch.randn(Cc, generator=g) * 0.1
    want = RI.swish(RI.conv_same(x, w, s, groups=Cc) + b.view(1, -1, 1, 1))
    ops.arena(torch.device("cuda")).reset()
    wp = w[:, 0].permute(1, 2, 0).reshape(k * k, Cc).contiguous().cuda()
    got, pool = ops.dwconv2d(_nhwc(x), wp, b.cuda(), k, s, ops.SS_ACT_SWISH, want_pool=True)
    assert tuple(got.shape) == (shape[0], 1, want.shape[2], want.shape[3], Cc)
    assert rel_err(_nchw(got), want) < 1e-5
    sums = want.double().sum((2, 3))
    assert rel_err(pool[..., 0].cpu(), sums) < 1e-5
    # no activation, no pooling, output into a channel slice of a wider buffer
    buf = torch.zeros((shape[0], 1, want.shape[2], want.shape[3], Cc + 32), device="cuda")
    got2, none = ops.dwconv2d(_nhwc(x), wp, b.cuda(), k, s, ops.SS_ACT_NONE, out=buf[..., :Cc])
    assert none is None and rel_err(_nchw(buf[..., :Cc]), RI.conv_same(x, w, s, groups=Cc) + b.view(1, -1, 1, 1)) < 1e-5
    assert float(buf[..., Cc:].abs().max()) == 0.0


def test_se_gate_matches_oracle(ops):
    g = torch.Generator().manual_seed(5)
    N, Cc, Sq, px = 2, 1344, 56, 24 * 80
    y = torch.randn(N, Cc, 6, 7, generator=g)
    w1, b1 = torch.randn(Sq, Cc, generator=g) * 0.05, torch.randn(Sq, generator=g) * 0.1
    w2, b2 = torch.randn(Cc, Sq, generator=g) * 0.2, torch.randn(Cc, generator=g) * 0.1
    pool = torch.zeros(N, Cc, 2, dtype=torch.float64)
    pool[..., 0] = y.double().sum((2, 3))
    pool[..., 1] = 123.0                                            # slot 1 must be ignored
    mean = y.mean((2, 3))
    want = torch.sigmoid(F.linear(RI.swish(F.linear(mean, w1, b1)), w2, b2))
    h = ops.se_fc(pool.cuda(), w1.cuda(), b1.cuda(), ops.SS_ACT_SWISH, 1.0 / 42)
    got = ops.se_fc(h, w2.cuda(), b2.cuda(), ops.SS_ACT_SIGMOID)
    assert rel_err(got.cpu(), want) < 1e-5


def test_swish_epilogue_and_se_gate_as_pending_scale(ops):
    """expand (1x1 + Swish) and linear (1x1 on the SE-gated input) as the tcgen05 GEMMs see them, 48 -> 64 padded channels."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(6)
    N, H, W = 2, 24, 40
    x = torch.randn(N, 48, H, W, generator=g)
    conv = nn.Conv2d(64, 288, 1, bias=True)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(288, 64, 1, 1, generator=g) * 0.15)
        conv.weight[:, 48:] = 0
        conv.bias.copy_(torch.randn(288, generator=g) * 0.1)
    want = RI.swish(F.conv2d(x, conv.weight[:, :48], conv.bias))
    xb = torch.zeros((N, 1, H, W, 64), device="cuda")
    xb[..., :48] = _nhwc(x)
    for mode, tol in ((ops.SS_MATH_TF32X3, 2e-5), (ops.SS_MATH_TF32, 2e-3), (ops.SS_MATH_3XTF32, 2e-5)):
        got, _ = ops.conv(ops.Vol(xb), conv.cuda(), out_act=ops.SS_ACT_SWISH, math_mode=mode)
        assert rel_err(_nchw(got), want) < tol, mode
    lin = nn.Conv2d(288, 48, 1, bias=True)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(48, 288, 1, 1, generator=g) * 0.1)
    gate = torch.rand(N, 288, generator=g)
    want2 = F.conv2d(want * gate.view(N, 288, 1, 1), lin.weight, lin.bias).detach()
    out = torch.zeros((N, 1, H, W, 64), device="cuda")
    z = _nhwc(want)
    for mode, tol in ((ops.SS_MATH_TF32X3, 2e-5), (ops.SS_MATH_TF32, 2e-3)):
        ops.conv(ops.Vol(z, gate.cuda(), torch.zeros(N, 288, device="cuda")), lin.cuda(), out=out[..., :48], math_mode=mode)
        assert rel_err(_nchw(out[..., :48]), want2) < tol, mode
        assert float(out[..., 48:].abs().max()) == 0.0
    # identity shortcut taken in the GEMM epilogue, in place on the (channel-padded) block input
    for mode, tol in ((ops.SS_MATH_TF32X3, 2e-5), (ops.SS_MATH_TF32, 2e-3), (ops.SS_MATH_3XTF32, 2e-5)):
        res = xb.clone()
        ops.conv(ops.Vol(z, gate.cuda(), torch.zeros(N, 288, device="cuda")), lin.cuda(), out=res[..., :48], math_mode=mode,
                 accumulate=True)
        assert rel_err(_nchw(res[..., :48]), want2 + x) < tol, mode
        assert float(res[..., 48:].abs().max()) == 0.0
    ops.use_f16x3(False)                     # the three-launch compensated mode accumulates too
    try:
        res = xb.clone()
        ops.conv(ops.Vol(z, gate.cuda(), torch.zeros(N, 288, device="cuda")), lin.cuda(), out=res[..., :48],
                 math_mode=ops.SS_MATH_TF32X3, accumulate=True)
    finally:
        ops.use_f16x3(True)
    assert rel_err(_nchw(res[..., :48]), want2 + x) < 2e-5


@pytest.fixture(scope="module")
def tiny_encoder():
    meta, gold = golden_image("tiny")
    enc = build_image_encoder(meta["seed"], "cuda")
    return meta, gold, enc


def _run(enc, img):
    from stereoscene_b200 import ops
    ops.arena(img.device).reset()
    levels = enc["img_backbone"].forward_vol(img.flatten(0, 1))
    feat = enc["img_neck"].forward_vol(levels)
    st = {f"img_level{i}": _nchw(buf[..., :c]) for i, (buf, c) in zip(enc["img_backbone"].out_indices, levels)}
    st["img_feat"] = _nchw(feat)
    return st


@pytest.mark.parametrize("policy,tol", [("mixed", 1e-4), ("tf32", 5e-3), ("3xtf32", 1e-4)])
def test_image_encoder_matches_reference_golden_tiny(ops, tiny_encoder, policy, tol):
    meta, gold, enc = tiny_encoder
    ops.set_math_policy(policy)
    try:
        with torch.no_grad():
            st = _run(enc, image_inputs(meta, "cuda"))
            st2 = _run(enc, image_inputs(meta, "cuda"))                # persistent padded buffers: a second call must agree
    finally:
        ops.set_math_policy(None)
    for k in LEVELS:
        assert rel_err(st[k], gold[k]) < tol, (policy, k)
        assert torch.equal(st[k], st2[k]), k


def test_image_encoder_matches_reference_golden_full_size(ops):
    """384x1280 stereo pair (the input size of BASELINE.json configs[1..4]) under the product policy, against the reference's
    own forward (strided samples, errors normalised by the reference's full-tensor statistics) and the live CPU oracle."""
    meta, gold = golden_image("full")
    enc = build_image_encoder(meta["seed"], "cuda")
    img = image_inputs(meta, "cuda")
    with torch.no_grad():
        st = _run(enc, img)
    report = {}
    for k in LEVELS:
        assert list(st[k].shape) == meta["stats"][k]["shape"]
        sl = tuple(slice(*s) for s in meta["samplers"][k])
        report[k] = stage_error(st[k][sl], gold[k], meta["stats"][k])
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_image_encoder_full.json", "w") as f:
        json.dump(report, f, indent=1)
    for k, e in report.items():
        assert e["max"] < 1e-4 and e["rms"] < 1e-4, (k, e)
    with torch.no_grad():
        want = RI.image_encoder(cpu_state_dict(enc), img.cpu())
    assert rel_err(st["img_feat"], want.flatten(0, 1)) < 1e-4


def test_reference_tensor_contract_and_end_to_end_from_images(ops):
    """CustomEfficientNet.forward / SECONDFPN.forward keep the reference's NCHW contract, and the detector run from images
    equals the detector run from those images' features (the channels-last hand-over changes no value)."""
    from stereoscene_b200 import presets, synth
    model, mc = presets.build("config0", image_encoder=True)
    synth.randomize_weights_(model, 4)
    model = model.cuda().eval()
    left_img, right_img = synth.stereo_images(1, mc["input_size"], seed=4, device="cuda")
    left, right, calib = synth.kitti_calibration(1, mc["input_size"], device="cuda")
    with torch.no_grad():
        levels = model.img_backbone(left_img.flatten(0, 1))
        assert [tuple(t.shape[1:]) for t in levels] == [(48, 32, 64), (80, 16, 32), (224, 8, 16), (640, 4, 8), (2560, 4, 8)]
        feat = model.img_neck([t.contiguous() for t in levels])[0]
        want = RI.image_encoder(cpu_state_dict(model), left_img.cpu())
        assert rel_err(feat.cpu(), want.flatten(0, 1)) < 1e-4
        enc = model.image_encoder(torch.cat([left_img, right_img], 0))
        assert tuple(enc.shape) == (2, 1, 640, 16, 32) and rel_err(enc[0, 0].cpu(), want[0, 0]) < 1e-4
        a = model.forward_images(left_img, right_img, left, right, calib, occ_size=mc["occ_size"], want_labels=True)
        b = model.forward_features(enc[:1].contiguous(), enc[1:].contiguous(), left, right, calib, occ_size=mc["occ_size"],
                                   want_labels=True)
    assert torch.equal(a["output_voxels"], b["output_voxels"]) and torch.equal(a["labels"], b["labels"])
    assert torch.isfinite(a["output_voxels"]).all()


@pytest.mark.parametrize("accumulate", [False, True])
def test_split_k_projection_matches_oracle(ops, accumulate):
    """Late-stage projection (960 pixels, K = 2304): 30 tiles on 148 SMs -> the box kernel splits K over CTAs when lent a workspace;
    the partial tiles are summed in a fixed order, so two runs agree bit for bit and the result equals the unsplit kernel's to rounding."""
    import torch.nn as nn
    from stereoscene_b200 import cabi
    g = torch.Generator().manual_seed(8)
    N, H, W, ci, co = 2, 12, 40, 2304, 384
    x = torch.randn(N, ci, H, W, generator=g)
    gate = torch.rand(N, ci, generator=g)
    lin = nn.Conv2d(ci, co, 1, bias=True)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(co, ci, 1, 1, generator=g) * 0.03)
    res = torch.randn(N, co, H, W, generator=g)
    want = F.conv2d(x.double() * gate.double().view(N, ci, 1, 1), lin.weight.double(), lin.bias.double()).float().detach()
    if accumulate:
        want = want + res
    ws = torch.empty(8 << 20, device="cuda")
    v = ops.Vol(_nhwc(x), gate.cuda(), torch.zeros(N, ci, device="cuda"))
    outs = []
    for use_ws in (ws, ws, None):
        out = _nhwc(res) if accumulate else torch.zeros((N, 1, H, W, co), device="cuda")
        c0 = cabi.kernel_census().get("splitk_reduce_kernel", 0)
        ops.conv(v, lin.cuda(), out=out, math_mode=ops.SS_MATH_TF32X3, accumulate=accumulate, splitk_ws=use_ws)
        assert cabi.kernel_census().get("splitk_reduce_kernel", 0) - c0 == (1 if use_ws is not None else 0)
        outs.append(_nchw(out))
    assert rel_err(outs[0], want) < 2e-5 and rel_err(outs[2], want) < 2e-5
    assert torch.equal(outs[0], outs[1])
    # a large-M layer ignores the workspace (its tiles already fill the GPU)
    big = nn.Conv2d(64, 96, 1, bias=True).cuda()
    xb = torch.randn(2, 1, 96, 320, 64, device="cuda")
    c0 = cabi.kernel_census().get("splitk_reduce_kernel", 0)
    ops.conv(ops.Vol(xb), big, math_mode=ops.SS_MATH_TF32X3, splitk_ws=ws)
    assert cabi.kernel_census().get("splitk_reduce_kernel", 0) == c0
