"""``CustomResNet3D`` and ``SECONDFPN3D`` -- B200-native replacements (same registry names, ctor
kwargs, forward contracts and state_dict keys as projects/mmdet3d_plugin/occupancy/backbones/
resnet3d.py:106-246 and necks/second_fpn_3d.py:13-117).

Every convolution is ss_conv3d_fwd with the previous layer's GroupNorm(+ReLU) applied as a
pending affine on load and this layer's GroupNorm sums accumulated in the epilogue; the only
elementwise pass per BasicBlock is the residual join.  The neck's three deblocks write straight
into channel slices of one [B,X,Y,Z,384] buffer (no torch.cat) and hand their GroupNorm+ReLU to
the head as a pending affine (``forward_vol``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import SS_ACT_NONE, SS_ACT_RELU, Vol
from ..registry import BACKBONES, NECKS
from .layers import as_channels_last


def _norm(norm_cfg, c) -> nn.Module:
    cfg = dict(norm_cfg)
    t = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if t != "GN":
        raise NotImplementedError(f"norm type {t}: stereoscene.py uses GroupNorm (norm_cfg type='GN') in the 3-D encoder")
    return nn.GroupNorm(cfg.pop("num_groups"), c, **cfg)


class BasicBlock3dParams(nn.Module):
    """keys: conv1, bn1, conv2, bn2, downsample.{0,1} (resnet3d.py:35-65, 196-198)."""

    def __init__(self, cin, planes, stride, norm_cfg):
        super().__init__()
        self.conv1 = nn.Conv3d(cin, planes, 3, stride, 1, bias=False)
        self.bn1 = _norm(norm_cfg, planes)
        self.conv2 = nn.Conv3d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = _norm(norm_cfg, planes)
        self.downsample = None
        if stride != 1 or cin != planes:
            self.downsample = nn.Sequential(nn.Conv3d(cin, planes, 1, stride, bias=False), _norm(norm_cfg, planes))
        self.stride = stride

    def run(self, x: Vol) -> torch.Tensor:
        y, st = ops.conv(x, self.conv1, want_stats=True)
        v = ops.gn_pending(y, st, self.bn1, SS_ACT_RELU)
        y, st = ops.conv(v, self.conv2, want_stats=True)
        v = ops.gn_pending(y, st, self.bn2, SS_ACT_NONE)
        if self.downsample is not None:
            r, st = ops.conv(x, self.downsample[0], want_stats=True)
            x = ops.gn_pending(r, st, self.downsample[1], SS_ACT_NONE)
        return ops.join(v, x, out_act=SS_ACT_RELU)


@BACKBONES.register_module()
class CustomResNet3D(nn.Module):
    _LAYERS = {10: [1, 1, 1, 1], 18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}

    def __init__(self, depth, block_inplanes=(64, 128, 256, 512), block_strides=(1, 2, 2, 2), out_indices=(0, 1, 2, 3),
                 num_stage=4, n_input_channels=3, shortcut_type="B", norm_cfg=dict(type="BN3d", requires_grad=True),
                 crp3d=False, crp_level=2, widen_factor=1.0, init_cfg=None):
        super().__init__()
        if depth not in self._LAYERS:
            raise NotImplementedError("only BasicBlock depths (10/18/34) are on the stereoscene.py path")
        if crp3d or shortcut_type != "B":
            raise NotImplementedError("crp3d / shortcut type A are not used by stereoscene.py")
        planes = [int(c * widen_factor) for c in block_inplanes]
        self.out_indices = tuple(out_indices)
        self.num_stage = num_stage
        self.crp3d = crp3d
        cin = planes[0]
        self.input_proj = nn.Sequential(nn.Conv3d(n_input_channels, cin, 1, 1, bias=False), _norm(norm_cfg, cin),
                                        nn.ReLU(inplace=True))
        self.layers = nn.ModuleList()
        for i, (c, n) in enumerate(zip(planes, self._LAYERS[depth])):
            if i + 1 > num_stage:
                break
            blocks = [BasicBlock3dParams(cin, c, block_strides[i], norm_cfg)]
            cin = c
            blocks += [BasicBlock3dParams(c, c, 1, norm_cfg) for _ in range(1, n)]
            self.layers.append(nn.Sequential(*blocks))
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self.forward_dic = {}

    def forward_vol(self, x: Vol):
        """x: pending/plain volume [B,X,Y,Z,C] -> list of plain channels-last tensors per stage."""
        y, st = ops.conv(x, self.input_proj[0], want_stats=True)
        v = ops.gn_pending(y, st, self.input_proj[1], SS_ACT_RELU)
        res = []
        for i, layer in enumerate(self.layers):
            with ops.math_scope(f"voxel.stage{i}"):          # per-stage math policy entry (falls back to "voxel")
                for blk in layer:
                    v = Vol(blk.run(v))
            if i in self.out_indices:
                res.append(v.data)
        return res

    def forward(self, x: torch.Tensor):
        """x: logical [B,C,X,Y,Z] -> list of logical [B,C_i,X_i,Y_i,Z_i] (resnet3d.py:219-246)."""
        ops.arena(x.device).reset()
        return [t.permute(0, 4, 1, 2, 3) for t in self.forward_vol(Vol(as_channels_last(x)))]


@NECKS.register_module()
class SECONDFPN3D(nn.Module):
    def __init__(self, in_channels=(128, 128, 256), out_channels=(256, 256, 256), upsample_strides=(1, 2, 4),
                 norm_cfg=dict(type="GN", num_groups=32, requires_grad=True), upsample_cfg=dict(type="deconv3d", bias=False),
                 conv_cfg=dict(type="Conv3d", bias=False), use_conv_for_no_stride=False, use_output_upsample=False,
                 with_cp=False, init_cfg=None):
        super().__init__()
        if use_output_upsample or use_conv_for_no_stride:
            raise NotImplementedError("use_output_upsample / use_conv_for_no_stride are not used by stereoscene.py")
        if upsample_cfg.get("type") != "deconv3d" or any(int(s) != s or s < 1 for s in upsample_strides):
            raise NotImplementedError("SECONDFPN3D on this path uses integer-stride deconv3d levels")
        assert len(out_channels) == len(upsample_strides) == len(in_channels)
        self.in_channels, self.out_channels = list(in_channels), list(out_channels)
        bias = bool(upsample_cfg.get("bias", True))
        self.deblocks = nn.ModuleList([
            nn.Sequential(nn.ConvTranspose3d(ci, co, int(s), int(s), bias=bias), _norm(norm_cfg, co), nn.ReLU(inplace=True))
            for ci, co, s in zip(in_channels, out_channels, upsample_strides)])

    def forward_vol(self, xs) -> Vol:
        """xs: list of channels-last tensors -> the concatenated output as ONE pending volume
        (raw deconv outputs side by side + per-level GroupNorm/ReLU as pending affine)."""
        assert len(xs) == len(self.in_channels)
        x0 = xs[0]
        B = x0.shape[0]
        s0 = self.deblocks[0][0].stride[0]
        X, Y, Z = x0.shape[1] * s0, x0.shape[2] * s0, x0.shape[3] * s0
        ctot = sum(self.out_channels)
        buf = torch.empty((B, X, Y, Z, ctot), dtype=torch.float32, device=x0.device)
        ss = torch.empty((2, B, ctot), dtype=torch.float32, device=x0.device)
        off = 0
        for x, blk, co in zip(xs, self.deblocks, self.out_channels):
            y, st = ops.conv(Vol(x), blk[0], out=buf[..., off:off + co], want_stats=True)
            ops.gn_pending(y, st, blk[1], SS_ACT_RELU, ss[0][:, off:off + co], ss[1][:, off:off + co])
            off += co
        return Vol(buf, ss[0], ss[1], SS_ACT_RELU)

    def forward(self, x):
        """x: list of logical [B,C_i,...] -> [logical [B,sum(C_out),X,Y,Z]] (second_fpn_3d.py:97-117)."""
        ops.arena(x[0].device).reset()
        v = self.forward_vol([as_channels_last(t) for t in x])
        return [v.ncdhw()]
