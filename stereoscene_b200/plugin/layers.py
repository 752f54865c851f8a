"""Parameter containers and fused layer drivers shared by the plugin modules.

The nn.Module trees below exist to hold parameters under the reference's state_dict keys
(SURVEY.md section 8b), so the published checkpoint loads with strict=True.  Their ``forward``
is never the nn.Sequential default: the drivers in this file walk the containers and issue the
C-ABI kernels (stereoscene_b200.ops), passing GroupNorm / BatchNorm / gates on as pending
affines.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import SS_ACT_GELU, SS_ACT_NONE, SS_ACT_RELU, Vol


# ------------------------------------------------------------------------------------------
# containers (keys follow ViewTransformerLSSVoxel.py:66-96 and attention.py:45-120)
# ------------------------------------------------------------------------------------------
def conv_gn3d(cin, cout, k, stride, pad, groups=2) -> nn.Sequential:
    """keys: 0.weight, 1.weight, 1.bias  (the reference's ``convbn_3d``: Conv3d(no bias)+GroupNorm(2))."""
    return nn.Sequential(nn.Conv3d(cin, cout, k, stride, pad, bias=False), nn.GroupNorm(groups, cout))


class HourglassParams(nn.Module):
    """keys: conv{1..4}.0.{0,1}, conv{5,6}.{0,1}, redir{1,2}.{0,1}  (ViewTransformerLSSVoxel.py:70-88)."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Sequential(conv_gn3d(c, 2 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(conv_gn3d(2 * c, 2 * c, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv3 = nn.Sequential(conv_gn3d(2 * c, 4 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(conv_gn3d(4 * c, 4 * c, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv5 = nn.Sequential(nn.ConvTranspose3d(4 * c, 2 * c, 3, padding=1, output_padding=1, stride=2, bias=False),
                                   nn.BatchNorm3d(2 * c))
        self.conv6 = nn.Sequential(nn.ConvTranspose3d(2 * c, c, 3, padding=1, output_padding=1, stride=2, bias=False),
                                   nn.BatchNorm3d(c))
        self.redir1 = conv_gn3d(c, c, 1, 1, 0)
        self.redir2 = conv_gn3d(2 * c, 2 * c, 1, 1, 0)


class MlpParams(nn.Module):
    """keys: fc1, fc2 (ViewTransformerLSSBEVDepth.py:417-439)."""

    def __init__(self, cin, hidden, cout):
        super().__init__()
        self.fc1 = nn.Linear(cin, hidden)
        self.fc2 = nn.Linear(hidden, cout)

    def forward(self, x):
        return self.fc2(torch.relu(self.fc1(x)))


class SEParams(nn.Module):
    """keys: conv_reduce, conv_expand (ViewTransformerLSSBEVDepth.py:442-454)."""

    def __init__(self, c):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, c, 1, bias=True)
        self.conv_expand = nn.Conv2d(c, c, 1, bias=True)

    def gate(self, x_se):
        """sigmoid(expand(relu(reduce(x_se)))) for x_se [B,C]: a per-(batch,channel) vector."""
        w1 = self.conv_reduce.weight.flatten(1)
        w2 = self.conv_expand.weight.flatten(1)
        h = torch.relu(torch.addmm(self.conv_reduce.bias, x_se, w1.t()))
        return torch.sigmoid(torch.addmm(self.conv_expand.bias, h, w2.t()))

    def forward(self, x, x_se):
        return x * self.gate(x_se.flatten(1))[..., None, None]


# ------------------------------------------------------------------------------------------
# fused drivers
# ------------------------------------------------------------------------------------------
def conv_gn(x: Vol, seq: nn.Sequential, act: int, out=None, scale_out=None, shift_out=None) -> Vol:
    """Conv -> (sums in the epilogue) -> pending GroupNorm + activation."""
    y, st = ops.conv(x, seq[0], out=out, want_stats=True)
    return ops.gn_pending(y, st, seq[1], act, scale_out, shift_out)


def hourglass(hg: HourglassParams, x: Vol) -> torch.Tensor:
    """ViewTransformerLSSVoxel.py:89-96.  11 convolutions, no normalisation pass; the two residual joins run in the
    epilogues of the up-convolutions where the layer shapes qualify (ops.conv_join), else as join kernels."""
    c1 = conv_gn(x, hg.conv1[0], SS_ACT_RELU)
    c2 = conv_gn(c1, hg.conv2[0], SS_ACT_RELU)
    c3 = conv_gn(c2, hg.conv3[0], SS_ACT_RELU)
    c4 = conv_gn(c3, hg.conv4[0], SS_ACT_RELU)
    r2 = conv_gn(c2, hg.redir2, SS_ACT_NONE)
    c5 = ops.conv_join(c4, hg.conv5[0], ops.bn_pending(c4.data, hg.conv5[1]), r2, out_act=SS_ACT_RELU)
    r1 = conv_gn(x, hg.redir1, SS_ACT_NONE)
    return ops.conv_join(Vol(c5), hg.conv6[0], ops.bn_pending(c5, hg.conv6[1]), r1, out_act=SS_ACT_RELU)


def as_channels_last(x: torch.Tensor) -> torch.Tensor:
    """Logical [B,C,D,H,W] -> [B,D,H,W,C] contiguous storage; free if x already is channels-last
    memory (which every tensor produced by this package is)."""
    v = x.permute(0, 2, 3, 4, 1)
    if v.is_contiguous():
        return v
    return ops.to_channels_last(x)
