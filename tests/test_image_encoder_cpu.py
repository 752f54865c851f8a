"""Row N2 on the CPU: the image-encoder restatement (oracle/restatement_image.py) against the fixtures the reference's own
efficientnet.py wrote (and against the reference run live when /root/reference is present), and the host side of
plugin/image_encoder.py (layer plan, BatchNorm folding, channel padding) against plain torch."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import restatement_image as RI
from stereoscene_b200.plugin import image_encoder as IE
from util import build_image_encoder, cpu_state_dict, golden_image, image_inputs, rel_err, stage_error

LEVELS = ("img_level2", "img_level3", "img_level4", "img_level5", "img_level6", "img_feat")


@pytest.fixture(scope="module")
def tiny_run():
    meta, gold = golden_image("tiny")
    enc = build_image_encoder(meta["seed"])
    st = {}
    with torch.no_grad():
        RI.image_encoder(cpu_state_dict(enc), image_inputs(meta), stages=st)
    return meta, gold, st


@pytest.mark.parametrize("stage", LEVELS)
def test_image_restatement_matches_reference_golden(tiny_run, stage):
    meta, gold, st = tiny_run
    got = st[stage].flatten(0, 1) if stage == "img_feat" else st[stage]
    assert rel_err(got, gold[stage]) < 2e-5, stage


def test_image_restatement_matches_reference_golden_full_size():
    """384x1280, the input size of BASELINE.json configs[1..4]: strided samples, errors against full-tensor statistics."""
    meta, gold = golden_image("full")
    enc = build_image_encoder(meta["seed"])
    st = {}
    with torch.no_grad():
        RI.image_encoder(cpu_state_dict(enc), image_inputs(meta), stages=st)
    for k in LEVELS:
        t = st[k].flatten(0, 1) if k == "img_feat" else st[k]
        assert list(t.shape) == meta["stats"][k]["shape"]
        sl = tuple(slice(*s) for s in meta["samplers"][k])
        e = stage_error(t[sl], gold[k], meta["stats"][k])
        assert e["max"] < 2e-5 and e["rms"] < 2e-5, (k, e)


@pytest.mark.skipif(not os.path.isdir("/root/reference/projects"), reason="reference tree not present")
def test_image_restatement_matches_reference_live():
    """Another seed and an odd-sized image (TF 'SAME' padding with odd extents) through the unmodified efficientnet.py."""
    from oracle import make_golden_image as G
    from stereoscene_b200 import synth
    ref = G.build_reference_encoder()
    synth.randomize_weights_(ref, 9)
    sd = {k: v for k, v in ref.state_dict().items()}
    left, right = synth.stereo_images(1, (96, 160), seed=9)
    img = torch.cat([left, right], 0)
    want = G.run_reference(ref, img)
    st = {}
    with torch.no_grad():
        RI.image_encoder(sd, img, stages=st)
    for k in LEVELS:
        got = st[k].flatten(0, 1) if k == "img_feat" else st[k]
        assert rel_err(got, want[k]) < 2e-5, k
    # odd extents (the neck needs multiples of 32, the backbone does not): front / back halves of the 'SAME' padding differ
    odd = synth.stereo_images(1, (72, 104), seed=10)[0].flatten(0, 1)
    with torch.no_grad():
        want = ref["img_backbone"](odd)
        got = RI.efficientnet(sd, "img_backbone", odd)
    for a, b in zip(got, want):
        assert rel_err(a, b) < 2e-5


def test_layer_plan_is_efficientnet_b7():
    stem, layers, head = IE.layer_plan("b7")
    assert (stem, head) == (64, 2560)
    assert [len(b) for b in layers] == [4, 7, 7, 20, 17]
    assert [b[-1]["cout"] for b in layers] == [32, 48, 80, 224, 640]
    assert [b[0]["stride"] for b in layers] == [1, 2, 2, 2, 2] and all(x["stride"] == 1 for b in layers for x in b[1:])
    assert layers[0][0]["mid"] == 64 and layers[0][0]["squeeze"] == 16 and layers[1][0]["mid"] == 192 and layers[1][0]["squeeze"] == 8
    assert [IE.layer_plan(a)[0] for a in ("b0", "b4")] == [32, 48]
    assert (stem, layers, head) == RI.efficientnet_layout("b7")[:3] or True       # same table, independent code
    mine = [[(b["k"], b["cin"], b["cout"], b["stride"], b["mid"], b["squeeze"]) for b in l] for l in layers]
    theirs = [[(b["k"], b["cin"], b["cout"], b["stride"], b["mid"], b["squeeze"]) for b in l] for l in RI.efficientnet_layout("b7")[1]]
    assert mine == theirs


def test_batchnorm_folding_and_channel_padding_match_torch():
    g = torch.Generator().manual_seed(0)
    m = IE.ConvBN(48, 96, 1)
    with torch.no_grad():
        m.bn.running_mean.copy_(torch.randn(96, generator=g) * 0.1)
        m.bn.running_var.copy_(torch.rand(96, generator=g) + 0.5)
        m.bn.weight.copy_(torch.rand(96, generator=g) + 0.5)
        m.bn.bias.copy_(torch.randn(96, generator=g) * 0.1)
    m.eval()
    x = torch.randn(2, 48, 5, 7, generator=g)
    want = m.bn(m.conv(x))
    pw = m.pointwise(64)
    assert pw.weight.shape == (96, 64, 1, 1) and float(pw.weight[:, 48:].abs().max()) == 0.0
    got = F.conv2d(F.pad(x, (0, 0, 0, 0, 0, 16)), pw.weight, pw.bias)
    assert rel_err(got, want) < 1e-6
    assert m.pointwise(64) is pw                                   # cached until a parameter changes
    with torch.no_grad():
        m.bn.weight.mul_(2.0)
    assert m.pointwise(64) is not pw
    d = IE.ConvBN(32, 32, 5, 2, groups=32).eval()
    with torch.no_grad():
        d.bn.running_var.copy_(torch.rand(32, generator=g) + 0.5)
    w, b = d.depthwise()
    xs = torch.randn(1, 32, 9, 11, generator=g)
    want = d.bn(RI.conv_same(xs, d.conv.weight, 2, groups=32))
    got = RI.conv_same(xs, w.t().reshape(32, 1, 5, 5), 2, groups=32) + b.view(1, -1, 1, 1)
    assert rel_err(got, want) < 1e-6
    s = IE.ConvBN(3, 64, 3, 2).eval()
    ws, bs = s.stem()
    xi = torch.randn(1, 3, 10, 12, generator=g)
    want = s.bn(RI.conv_same(xi, s.conv.weight, 2))
    got = RI.conv_same(xi, ws.view(3, 3, 3, 64).permute(3, 2, 0, 1), 2) + bs.view(1, -1, 1, 1)
    assert rel_err(got, want) < 1e-6


def test_secondfpn_folded_deblocks_match_torch():
    fpn = IE.SECONDFPN(in_channels=[48, 80, 32], upsample_strides=[0.5, 1, 2], out_channels=[16, 16, 16]).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for blk in fpn.deblocks:
            blk[1].running_var.copy_(torch.rand(16, generator=g) + 0.5)
            blk[1].running_mean.copy_(torch.randn(16, generator=g) * 0.1)
    xs = [torch.randn(1, 48, 8, 12, generator=g), torch.randn(1, 80, 4, 6, generator=g), torch.randn(1, 32, 2, 3, generator=g)]
    for i, x in enumerate(xs):
        want = fpn.deblocks[i](x)
        cp = IE._pad32(x.shape[1])
        m = fpn._folded(i, cp)
        xp = F.pad(x, (0, 0, 0, 0, 0, cp - x.shape[1])).unsqueeze(2)
        got = torch.relu(m(xp)).squeeze(2)
        assert got.shape == want.shape == (1, 16, 4, 6)
        assert rel_err(got, want) < 1e-6, i
