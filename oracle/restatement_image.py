"""TEST INFRASTRUCTURE ONLY -- CPU restatement (PyTorch, functional) of the reference's 2-D image encoder,
SURVEY.md section 8 row N2: ``CustomEfficientNet`` arch 'b7' (projects/mmdet3d_plugin/occupancy/backbones/
efficientnet.py:274-534 = EFF) followed by mmdet3d's ``SECONDFPN`` (stereoscene.py:59-74 = CFG), as
``BEVDepthOccupancy.image_encoder`` calls them (detectors/bevdepth_occupancy.py:42-59 = DET).
Never imported by the product package ``stereoscene_b200``.

Third-party code the reference reaches here and that is absent from /root/reference (versions pinned by
docs/install.md) is restated from the published algorithms:
  * mmcv 1.4.0 ``ConvModule`` (conv -> norm -> act, no conv bias when a norm follows), ``Conv2dAdaptivePadding``
    (TensorFlow 'SAME' padding: total = max((ceil(H/s)-1)*s + k - H, 0), smaller half in front), ``Swish``;
  * mmdet 2.14.0 ``SELayer`` (global mean -> 1x1 conv + bias, act -> 1x1 conv + bias, sigmoid -> gate) and
    ``make_divisible``;
  * mmdet3d 0.17.1 ``SECONDFPN`` (per level: ConvTranspose2d k = s for s >= 1, Conv2d k = s = 1/s below 1, no bias,
    BatchNorm2d eps 1e-3, ReLU; concatenation over channels).

Pinned: tests/test_oracle_golden.py compares it with tests/golden/golden_image_*.npz, written by
oracle/make_golden_image.py from the reference's own efficientnet.py loaded through oracle/ref_loader.py
(on the same restated mmcv bricks -- the mmcv layer itself is therefore pinned to its published behaviour, not
to a run of mmcv; the architecture, the scaling rule and the block wiring are the reference file's own).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # EFF:373 norm_cfg=dict(type='BN', eps=1e-3); SECONDFPN default eps=1e-3

# EfficientNet-B0 stages (EFF:300-317): kernel, channels, stride, expand ratio, repeats; se_ratio is 4 everywhere
_B0 = ((3, 16, 1, 1, 1), (3, 24, 2, 6, 2), (5, 40, 2, 6, 2), (3, 80, 2, 6, 3), (5, 112, 1, 6, 3), (5, 192, 2, 6, 4),
       (3, 320, 1, 6, 1))
_ARCH = {"b0": (1.0, 1.0), "b1": (1.0, 1.1), "b2": (1.1, 1.2), "b3": (1.2, 1.4), "b4": (1.4, 1.8), "b5": (1.6, 2.2),
         "b6": (1.8, 2.6), "b7": (2.0, 3.1), "b8": (2.2, 3.6)}          # EFF:343-355 (width, depth)


def round_channels(v, divisor=8):
    """mmdet ``make_divisible`` (min_ratio 0.9)."""
    n = max(divisor, int(v + divisor / 2) // divisor * divisor)
    return n + divisor if n < 0.9 * v else n


def efficientnet_layout(arch="b7"):
    """The layers of ``CustomEfficientNet`` as (stem_channels, groups, head_channels): compound scaling of the B0
    stages (``model_scaling``, EFF:232-271: channels x width rounded to 8, repeats = ceil(n x depth) with the last
    block of a stage repeated), a stride-1 stage merged into the layer before it (EFF:265-268).  Every block is
    dict(k, cin, cout, stride, expand, squeeze)."""
    wmul, dmul = _ARCH[arch]
    stem = round_channels(32 * wmul)
    cin = stem
    groups = []
    for si, (k, c, s, e, n) in enumerate(_B0):
        cout = round_channels(c * wmul)
        blocks = []
        for i in range(int(math.ceil(dmul * n))):
            mid = int(cin * e)
            blocks.append(dict(k=k, cin=cin, cout=cout, stride=s if i == 0 else 1, expand=e, mid=mid,
                               squeeze=int(mid / (e * 4))))            # EFF:449-457 SE ratio = expand * se_ratio
            cin = cout
        if s == 1 and si > 0:
            groups[-1] += blocks
        else:
            groups.append(blocks)
    return stem, groups, round_channels(1280 * wmul)


def swish(x):
    return x * torch.sigmoid(x)


def same_pad(size, k, s):
    total = max((math.ceil(size / s) - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv_same(x, w, stride, groups=1):
    (t, b), (l, r) = same_pad(x.shape[-2], w.shape[-2], stride), same_pad(x.shape[-1], w.shape[-1], stride)
    if t or b or l or r:
        x = F.pad(x, [l, r, t, b])
    return F.conv2d(x, w, None, stride, 0, 1, groups)


def bn_eval(sd, p, x):
    scale = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + BN_EPS)
    shift = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def conv_module(sd, p, x, stride=1, groups=1, act=True):
    y = bn_eval(sd, p + ".bn", conv_same(x, sd[p + ".conv.weight"], stride, groups))
    return swish(y) if act else y


def mbconv(sd, p, x, blk):
    """``InvertedResidual.forward`` (EFF:205-231): expand 1x1 (absent when mid == cin) -> depthwise k x k -> SE
    -> linear 1x1 (no activation) -> identity shortcut when stride 1 and cin == cout; drop-path is the identity
    in eval."""
    y = x
    if blk["mid"] != blk["cin"]:
        y = conv_module(sd, p + ".expand_conv", y)
    y = conv_module(sd, p + ".depthwise_conv", y, blk["stride"], groups=blk["mid"])
    g = y.mean((2, 3), keepdim=True)
    g = swish(F.conv2d(g, sd[p + ".se.conv1.conv.weight"], sd[p + ".se.conv1.conv.bias"]))
    g = torch.sigmoid(F.conv2d(g, sd[p + ".se.conv2.conv.weight"], sd[p + ".se.conv2.conv.bias"]))
    y = conv_module(sd, p + ".linear_conv", y * g, act=False)
    return x + y if (blk["stride"] == 1 and blk["cin"] == blk["cout"]) else y


def efficientnet(sd, p, x, arch="b7", out_indices=(2, 3, 4, 5, 6)):
    """``CustomEfficientNet.forward`` (EFF:505-513): layer 0 = stem conv 3x3 s2, layers 1..5 = MBConv groups,
    layer 6 = 1x1 head conv; returns the outputs of ``out_indices``."""
    _, groups, _ = efficientnet_layout(arch)
    outs = []
    x = conv_module(sd, p + ".layers.0", x, stride=2)
    if 0 in out_indices:
        outs.append(x)
    for li, blocks in enumerate(groups, start=1):
        for bi, blk in enumerate(blocks):
            x = mbconv(sd, f"{p}.layers.{li}.{bi}", x, blk)
        if li in out_indices:
            outs.append(x)
    li = len(groups) + 1
    if li <= max(out_indices):                              # EFF:407-417: the head conv exists only if it is an output
        x = conv_module(sd, f"{p}.layers.{li}", x)
        outs.append(x)
    return outs


def second_fpn(sd, p, feats, upsample_strides=(0.5, 1, 2, 4, 4)):
    """mmdet3d SECONDFPN (see the module docstring), CFG:70-74."""
    ups = []
    for i, (x, s) in enumerate(zip(feats, upsample_strides)):
        w = sd[f"{p}.deblocks.{i}.0.weight"]
        if s >= 1:
            y = F.conv_transpose2d(x, w, None, stride=int(s))
        else:
            y = F.conv2d(x, w, None, stride=int(round(1 / s)))
        ups.append(F.relu(bn_eval(sd, f"{p}.deblocks.{i}.1", y)))
    return torch.cat(ups, 1)


def image_encoder(sd, img, arch="b7", out_indices=(2, 3, 4, 5, 6), upsample_strides=(0.5, 1, 2, 4, 4), stages=None):
    """``BEVDepthOccupancy.image_encoder`` (DET:42-59): [B,N,3,H,W] -> [B,N,640,H/8,W/8].  ``stages`` (a dict)
    receives the backbone outputs ``img_level{i}`` and ``img_feat``."""
    B, N, Cc, H, W = img.shape
    levels = efficientnet(sd, "img_backbone", img.reshape(B * N, Cc, H, W), arch, out_indices)
    x = second_fpn(sd, "img_neck", levels, upsample_strides)
    x = x.view(B, N, *x.shape[1:])
    if stages is not None:
        for i, t in zip(out_indices, levels):
            stages[f"img_level{i}"] = t
        stages["img_feat"] = x
    return x
