"""The 2-D image encoder in front of the volumetric path: ``CustomEfficientNet``
(projects/mmdet3d_plugin/occupancy/backbones/efficientnet.py:274-534, arch 'b7' in stereoscene.py:59-69) and mmdet3d's
``SECONDFPN`` (stereoscene.py:70-74), as ``BEVDepthOccupancy.image_encoder`` chains them (bevdepth_occupancy.py:42-59).

The module trees hold the parameters under the reference's state_dict keys (``layers.{i}.{j}.expand_conv.conv.weight``,
``...depthwise_conv.bn.running_var``, ``...se.conv1.conv.bias``, ``deblocks.{i}.0.weight`` ...), so the published
checkpoint loads with strict=True.  The forward is not theirs:

  * activations are channels-last fp32 depth-1 volumes [N,1,H,W,C]; eval-mode BatchNorm is folded into the convolution
    that precedes it (weights x scale, shift as the bias), Swish runs in that convolution's epilogue;
  * the pointwise convolutions (expand, linear, head, SECONDFPN deblocks -- 99 % of the FLOPs) are GEMMs on the tcgen05
    kernels through ``ops.conv``, in the math mode of the policy group "image";
  * the depthwise convolution also produces the squeeze-excite block's pooled sums; the SE gate is never multiplied
    into the activation -- it is the pending per-(image, channel) scale of the linear convolution's input;
  * 48- and 80-channel block outputs live in 64- / 96-channel buffers whose padding channels stay zero (the GEMM kernels
    take K in 32-channel chunks); the consuming layers carry zero weight rows for them;
  * the SECONDFPN deblocks write straight into channel slices of the concatenated [N,1,H/8,W/8,640] output, which is the
    channels-last feature pair the view transformer starts from (no layout change in between).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from ..ops import SS_ACT_RELU, SS_ACT_SIGMOID, SS_ACT_SWISH, Vol
from ..registry import BACKBONES, NECKS

BN_EPS = 1e-3

# EfficientNet-B0 stages: kernel, channels, stride, expand ratio, repeats (squeeze-excite ratio 4 everywhere);
# compound scaling (width, depth) per architecture -- Tan & Le 2019, as tabulated in efficientnet.py:300-317, 343-355
_B0_STAGES = ((3, 16, 1, 1, 1), (3, 24, 2, 6, 2), (5, 40, 2, 6, 2), (3, 80, 2, 6, 3), (5, 112, 1, 6, 3), (5, 192, 2, 6, 4),
              (3, 320, 1, 6, 1))
_SCALING = {"b0": (1.0, 1.0), "b1": (1.0, 1.1), "b2": (1.1, 1.2), "b3": (1.2, 1.4), "b4": (1.4, 1.8), "b5": (1.6, 2.2),
            "b6": (1.8, 2.6), "b7": (2.0, 3.1), "b8": (2.2, 3.6)}


def _round8(v: float) -> int:
    n = max(8, int(v + 4) // 8 * 8)
    return n + 8 if n < 0.9 * v else n


def _pad32(c: int) -> int:
    return (c + 31) // 32 * 32


def layer_plan(arch: str):
    """(stem channels, [[block dict, ...] per layer 1..5], head channels): every stage's channels x width rounded to 8 and
    repeats = ceil(n x depth); a stride-1 stage shares the layer of the stage before it (efficientnet.py:232-271)."""
    wmul, dmul = _SCALING[arch]
    stem = _round8(32 * wmul)
    cin, layers = stem, []
    for si, (k, c, s, e, n) in enumerate(_B0_STAGES):
        cout, blocks = _round8(c * wmul), []
        for i in range(int(math.ceil(n * dmul))):
            mid = int(cin * e)
            blocks.append(dict(k=k, cin=cin, cout=cout, stride=s if i == 0 else 1, mid=mid, squeeze=int(mid / (e * 4))))
            cin = cout
        if s == 1 and si > 0:
            layers[-1] += blocks
        else:
            layers.append(blocks)
    return stem, layers, _round8(1280 * wmul)


def _key(*tensors):
    return tuple((t.data_ptr(), t._version, t.device) for t in tensors)


def _bn_fold(bn: nn.BatchNorm2d):
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
    return scale, bn.bias.detach().float() - bn.running_mean.float() * scale


class ConvBN(nn.Module):
    """mmcv ``ConvModule`` as efficientnet.py uses it: keys conv.weight, bn.{weight,bias,running_mean,running_var,
    num_batches_tracked}.  ``pointwise()`` / ``depthwise()`` / ``stem()`` return the BatchNorm-folded kernel operands, cached
    until a parameter changes."""

    def __init__(self, cin, cout, k, stride=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, 0, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=BN_EPS)
        self.k, self.stride = k, stride
        object.__setattr__(self, "_cache", {})

    def _cached(self, tag, fn):
        key = _key(self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
        hit = self._cache.get(tag)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, fn())
            self._cache[tag] = hit
        return hit[1]

    def pointwise(self, cin_padded=None) -> nn.Conv2d:
        """1x1 conv holder [Cout, Cin (zero-padded to cin_padded)] with bias, for ops.conv."""
        def build():
            w = self.conv.weight.detach().float()
            scale, shift = _bn_fold(self.bn)
            cout, cin = w.shape[0], w.shape[1]
            cp = cin_padded or cin
            m = nn.Conv2d(cp, cout, 1, bias=True).to(w.device)
            m.weight.zero_()
            m.weight[:, :cin] = w * scale.view(-1, 1, 1, 1)
            m.bias.copy_(shift)
            m.requires_grad_(False)
            return m
        return self._cached(("pw", cin_padded), build)

    def depthwise(self):
        """(w [k*k, C], bias [C])."""
        def build():
            w = self.conv.weight.detach().float()                          # [C,1,k,k]
            scale, shift = _bn_fold(self.bn)
            return (w[:, 0] * scale.view(-1, 1, 1)).permute(1, 2, 0).reshape(-1, w.shape[0]).contiguous(), shift.contiguous()
        return self._cached("dw", build)

    def stem(self):
        """(w [k*k*Cin, Cout] ordered (ky, kx, ci), bias [Cout])."""
        def build():
            w = self.conv.weight.detach().float()                          # [Cout,Cin,k,k]
            scale, shift = _bn_fold(self.bn)
            return (w * scale.view(-1, 1, 1, 1)).permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous(), shift.contiguous()
        return self._cached("stem", build)


class _Conv1x1Bias(nn.Module):
    """key: conv.{weight,bias} (ConvModule without a norm layer: the SE block's two convolutions)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=True)


class SELayer(nn.Module):
    """mmdet ``SELayer`` parameters (keys conv1.conv.*, conv2.conv.*); ``gate`` = sigmoid(W2 swish(W1 mean + b1) + b2)."""

    def __init__(self, channels, squeeze):
        super().__init__()
        self.conv1 = _Conv1x1Bias(channels, squeeze)
        self.conv2 = _Conv1x1Bias(squeeze, channels)

    def gate(self, pool: torch.Tensor, pixels: int) -> torch.Tensor:
        c1, c2 = self.conv1.conv, self.conv2.conv
        h = ops.se_fc(pool, c1.weight.detach().flatten(1), c1.bias.detach(), SS_ACT_SWISH, 1.0 / pixels)
        return ops.se_fc(h, c2.weight.detach().flatten(1), c2.bias.detach(), SS_ACT_SIGMOID)


class InvertedResidual(nn.Module):
    """efficientnet.py:113-231 (keys expand_conv, depthwise_conv, se, linear_conv)."""

    def __init__(self, k, cin, cout, stride, mid, squeeze):
        super().__init__()
        self.cin, self.cout, self.mid, self.k, self.stride = cin, cout, mid, k, stride
        if mid != cin:
            self.expand_conv = ConvBN(cin, mid, 1)
        self.depthwise_conv = ConvBN(mid, mid, k, stride, groups=mid)
        self.se = SELayer(mid, squeeze)
        self.linear_conv = ConvBN(mid, cout, 1)
        self.with_res_shortcut = stride == 1 and cin == cout

    def forward_vol(self, x: torch.Tensor, bufs) -> torch.Tensor:
        """x: [N,1,H,W,pad32(cin)] -> [N,1,H',W',pad32(cout)] (padding channels zero)."""
        y = x
        if self.mid != self.cin:
            y, _ = ops.conv(Vol(x), self.expand_conv.pointwise(x.shape[-1]), out_act=SS_ACT_SWISH)
        w, b = self.depthwise_conv.depthwise()
        z, pool = ops.dwconv2d(y, w, b, self.k, self.stride, SS_ACT_SWISH, want_pool=True)
        N, _, Ho, Wo, _ = z.shape
        gate = self.se.gate(pool, Ho * Wo)
        zg = Vol(z, gate, bufs.zeros_like(gate))
        ws = bufs.workspace(x.device)
        if self.with_res_shortcut:       # x += linear(z * gate): the shortcut is taken in the GEMM's epilogue, in place on the block input
            ops.conv(zg, self.linear_conv.pointwise(), out=x[..., :self.cout], accumulate=True, splitk_ws=ws)
            return x
        out = bufs.take(N, Ho, Wo, self.cout, avoid=x)
        ops.conv(zg, self.linear_conv.pointwise(), out=out[..., :self.cout], splitk_ws=ws)
        return out


class _Buffers:
    """Block outputs.  Channel counts that are multiples of 32 come from the caching allocator; 48 / 80-channel outputs need
    persistent zero-padded buffers (two per shape, used alternately: a block's input stays alive until its residual join)."""

    def __init__(self):
        self.padded = {}
        self.zero = {}

    def take(self, N, H, W, C, avoid=None) -> torch.Tensor:
        dev = avoid.device
        cp = _pad32(C)
        if cp == C:
            return torch.empty((N, 1, H, W, C), dtype=torch.float32, device=dev)
        pair = self.padded.get((N, H, W, cp, dev))
        if pair is None:
            pair = self.padded[(N, H, W, cp, dev)] = [torch.zeros((N, 1, H, W, cp), dtype=torch.float32, device=dev) for _ in range(2)]
        return pair[1] if avoid.data_ptr() == pair[0].data_ptr() else pair[0]

    def workspace(self, dev) -> torch.Tensor:
        """Split-K scratch for the long-K projections of the late stages (8 slices x 960 pixels x 640 channels fit)."""
        w = self.zero.get(("ws", dev))
        if w is None:
            w = self.zero[("ws", dev)] = torch.empty(8 << 20, dtype=torch.float32, device=dev)
        return w

    def zeros_like(self, t: torch.Tensor) -> torch.Tensor:
        key = (tuple(t.shape), t.device)
        z = self.zero.get(key)
        if z is None:
            z = self.zero[key] = torch.zeros_like(t)
        return z


@BACKBONES.register_module()
class CustomEfficientNet(nn.Module):
    """efficientnet.py:274-534.  ``forward`` keeps the reference's tensor contract ([N,3,H,W] -> tuple of [N,C,h,w], as
    channels_last views); ``forward_vol`` returns the padded channels-last buffers the neck consumes."""

    def __init__(self, arch="b0", drop_path_rate=0.0, out_indices=(6,), frozen_stages=0, conv_cfg=None, norm_cfg=None,
                 act_cfg=None, norm_eval=False, with_cp=False, init_cfg=None):
        super().__init__()
        if arch not in _SCALING:
            raise KeyError(f"CustomEfficientNet: arch {arch!r} is not one of {sorted(_SCALING)} (the EdgeTPU variants are not on this path)")
        stem, plan, head = layer_plan(arch)
        self.arch, self.out_indices = arch, tuple(out_indices)
        if any(i not in range(len(plan) + 2) for i in self.out_indices):
            raise ValueError(f"out_indices must be in range(0, {len(plan) + 2}), got {out_indices}")
        self.layers = nn.ModuleList([ConvBN(3, stem, 3, 2)])
        for li, blocks in enumerate(plan, start=1):
            if li > max(self.out_indices):
                break                                                  # efficientnet.py:431-433: unused layers are not built
            self.layers.append(nn.Sequential(*[InvertedResidual(**b) for b in blocks]))
        if len(self.layers) < max(self.out_indices) + 1:
            self.layers.append(ConvBN(plan[-1][-1]["cout"], head, 1))
        self.level_channels = [stem] + [b[-1]["cout"] for b in plan] + [head]
        object.__setattr__(self, "_bufs", _Buffers())

    def forward_vol(self, img: torch.Tensor):
        """img [N,3,H,W] -> list of (buffer [N,1,h,w,pad32(C)], C) for ``out_indices``."""
        outs = []
        with ops.math_scope("image"):
            w, b = self.layers[0].stem()
            x = ops.stem_conv2d(img, w, b, 3, 2, SS_ACT_SWISH)
            if 0 in self.out_indices:
                outs.append((x, x.shape[-1]))
            for li in range(1, len(self.layers)):
                layer = self.layers[li]
                if isinstance(layer, ConvBN):
                    x, _ = ops.conv(Vol(x), layer.pointwise(x.shape[-1]), out_act=SS_ACT_SWISH)
                else:
                    for blk in layer:
                        x = blk.forward_vol(x, self._bufs)
                if li in self.out_indices:
                    outs.append((x, self.level_channels[li]))
        return outs

    def forward(self, x):
        return tuple(buf[..., :c].squeeze(1).permute(0, 3, 1, 2) for buf, c in self.forward_vol(x))


@NECKS.register_module()
class SECONDFPN(nn.Module):
    """mmdet3d 0.17.1 ``SECONDFPN`` as configured at stereoscene.py:70-74 (keys deblocks.{i}.0.weight, deblocks.{i}.1.*):
    ConvTranspose2d(kernel = stride) for upsample strides >= 1, Conv2d(kernel = stride = 1/s) below 1, no bias, BatchNorm2d
    (eps 1e-3) + ReLU, concatenation over channels."""

    def __init__(self, in_channels=(128, 128, 256), out_channels=(256, 256, 256), upsample_strides=(1, 2, 4), norm_cfg=None,
                 upsample_cfg=None, conv_cfg=None, use_conv_for_no_stride=False, init_cfg=None):
        super().__init__()
        assert len(in_channels) == len(out_channels) == len(upsample_strides)
        norm_cfg = dict(norm_cfg or dict(type="BN", eps=1e-3, momentum=0.01))
        self.in_channels, self.out_channels, self.upsample_strides = list(in_channels), list(out_channels), list(upsample_strides)
        blocks = []
        for cin, cout, s in zip(in_channels, out_channels, upsample_strides):
            if s > 1 or (s == 1 and not use_conv_for_no_stride):
                up = nn.ConvTranspose2d(cin, cout, int(s), stride=int(s), bias=False)
            else:
                k = int(round(1 / s))
                up = nn.Conv2d(cin, cout, k, stride=k, bias=False)
            bn = nn.BatchNorm2d(cout, eps=norm_cfg.get("eps", 1e-3), momentum=norm_cfg.get("momentum", 0.01))
            blocks.append(nn.Sequential(up, bn, nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(blocks)
        object.__setattr__(self, "_cache", {})

    def _folded(self, i: int, cin_padded: int) -> nn.Module:
        """Deblock i as a depth-1 3-D (transposed) convolution holder with the BatchNorm folded in."""
        up, bn = self.deblocks[i][0], self.deblocks[i][1]
        key = _key(up.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var) + (cin_padded,)
        hit = self._cache.get(i)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                scale, shift = _bn_fold(bn)
                w = up.weight.detach().float()
                k = up.kernel_size[0]
                if isinstance(up, nn.ConvTranspose2d) and k == 1:      # a 1x1 transposed conv is a 1x1 conv on W^T
                    cin, cout = w.shape[0], w.shape[1]
                    m = nn.Conv3d(cin_padded, cout, 1, bias=True).to(w.device)
                    m.weight.zero_()
                    m.weight[:, :cin] = (w * scale.view(1, -1, 1, 1)).permute(1, 0, 2, 3).unsqueeze(2)
                elif isinstance(up, nn.ConvTranspose2d):               # [Cin,Cout,k,k]
                    cin, cout = w.shape[0], w.shape[1]
                    m = nn.ConvTranspose3d(cin_padded, cout, (1, k, k), stride=(1, k, k), bias=True).to(w.device)
                    m.weight.zero_()
                    m.weight[:cin] = (w * scale.view(1, -1, 1, 1)).unsqueeze(2)
                else:                                                  # [Cout,Cin,k,k]
                    cout, cin = w.shape[0], w.shape[1]
                    m = nn.Conv3d(cin_padded, cout, (1, k, k), stride=(1, k, k), bias=True).to(w.device)
                    m.weight.zero_()
                    m.weight[:, :cin] = (w * scale.view(-1, 1, 1, 1)).unsqueeze(2)
                m.bias.copy_(shift)
                m.requires_grad_(False)
            hit = (key, m)
            self._cache[i] = hit
        return hit[1]

    def forward_vol(self, levels) -> torch.Tensor:
        """levels: [(buffer [N,1,h,w,pad32(C)], C)] -> channels-last [N,1,H,W,sum(out_channels)]."""
        assert len(levels) == len(self.in_channels)
        out = None
        c0 = 0
        with ops.math_scope("image"):
            for i, ((buf, cc), s) in enumerate(zip(levels, self.upsample_strides)):
                assert cc == self.in_channels[i], (cc, self.in_channels[i])
                N, _, h, w, _ = buf.shape
                H, W = (int(h * s), int(w * s)) if s >= 1 else (h // int(round(1 / s)), w // int(round(1 / s)))
                if out is None:
                    out = torch.empty((N, 1, H, W, sum(self.out_channels)), dtype=torch.float32, device=buf.device)
                ops.conv(Vol(buf), self._folded(i, buf.shape[-1]), out=out[..., c0:c0 + self.out_channels[i]], out_act=SS_ACT_RELU)
                c0 += self.out_channels[i]
        return out

    def forward(self, x):
        levels = [(ops.to_channels_last(t.contiguous()).unsqueeze(1), t.shape[1]) for t in x]
        levels = [(b if b.shape[-1] % 32 == 0 else torch.nn.functional.pad(b, (0, _pad32(b.shape[-1]) - b.shape[-1])), c) for b, c in levels]
        return [self.forward_vol(levels).squeeze(1).permute(0, 3, 1, 2)]
