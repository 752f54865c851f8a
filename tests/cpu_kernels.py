"""TEST INFRASTRUCTURE ONLY -- a torch-CPU (float64) re-statement of the handful of ``stereoscene_b200.ops`` entry points
that ``stereoscene_b200.xshard`` calls, with the same Vol / pending-affine semantics, so the gloo tests can drive the
sharded orchestration (slab bookkeeping, halo exchange, statistics all-reduce) without a GPU.  Nothing in the product
imports this module."""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

SS_ACT_NONE, SS_ACT_RELU, SS_ACT_GELU = 0, 1, 2


@dataclass
class Vol:
    data: torch.Tensor
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    act: int = SS_ACT_NONE


def _act(t, act):
    return F.relu(t) if act == SS_ACT_RELU else (F.gelu(t) if act == SS_ACT_GELU else t)


def _value(v: Vol) -> torch.Tensor:
    """Logical value, NCDHW float64."""
    t = v.data.double().permute(0, 4, 1, 2, 3)
    if v.scale is not None:
        t = t * v.scale.double()[:, :, None, None, None] + v.shift.double()[:, :, None, None, None]
    return _act(t, v.act)


def conv(x: Vol, module, out=None, out_act=SS_ACT_NONE, want_stats=False, math_mode=None, use_bias=True, pad=None,
         stats_planes=None):
    t = _value(x)
    w = module.weight.detach().double()
    b = module.bias.detach().double() if (use_bias and module.bias is not None) else None
    if isinstance(module, torch.nn.ConvTranspose3d):
        y = F.conv_transpose3d(t, w, b, stride=module.stride, padding=module.padding if pad is None else tuple(pad),
                               output_padding=module.output_padding)
    else:
        y = F.conv3d(t, w, b, stride=module.stride, padding=module.padding if pad is None else tuple(pad),
                     dilation=module.dilation)
    y = _act(y, out_act)
    stats = None
    if want_stats:
        ys = y if stats_planes is None else y[:, :, stats_planes[0]:stats_planes[1]]
        stats = torch.stack([ys.sum(dim=(2, 3, 4)), (ys * ys).sum(dim=(2, 3, 4))], dim=-1).contiguous()
    ycl = y.permute(0, 2, 3, 4, 1)
    if out is None:
        out = ycl.contiguous().to(x.data.dtype)
    else:
        out.copy_(ycl)
    return out, stats


def gn_pending(y, stats, gn, act=SS_ACT_NONE, scale_out=None, shift_out=None, count=None):
    B, C = stats.shape[0], stats.shape[1]
    n = float(y.shape[1] * y.shape[2] * y.shape[3]) if count is None else float(count)
    G = gn.num_groups
    s = stats[..., 0].view(B, G, C // G).sum(-1)
    q = stats[..., 1].view(B, G, C // G).sum(-1)
    cnt = n * (C // G)
    mean = s / cnt
    var = q / cnt - mean * mean
    rstd = 1.0 / torch.sqrt(var + gn.eps)
    scale = gn.weight.detach().double()[None] * rstd.repeat_interleave(C // G, dim=1)
    shift = gn.bias.detach().double()[None] - mean.repeat_interleave(C // G, dim=1) * scale
    if scale_out is not None:
        scale_out.copy_(scale)
        shift_out.copy_(shift)
        scale, shift = scale_out, shift_out
    return Vol(y, scale, shift, act)


def join(x: Vol, r: Optional[Vol], out_act=SS_ACT_NONE, alpha=None, out=None):
    t = _value(x)
    if alpha is not None:
        t = t * alpha.double()
    if r is not None:
        t = t + _value(r)
    res = _act(t, out_act).permute(0, 2, 3, 4, 1).contiguous().to(x.data.dtype)
    if out is not None:
        out.copy_(res)
        return out
    return res


@dataclass
class SplatIndex:
    order: torch.Tensor
    voxel_start: torch.Tensor
    coords: Optional[torch.Tensor]
    nx: int
    ny: int
    nz: int
    B: int
    P: int


def splat_build_index(geom, dx, bx, nx):
    """CPU twin of ss_splat_build_index for one sample (points sorted by voxel rank, dropped points last)."""
    from oracle import restatement as O
    idx, kept = O.voxel_indices(geom.reshape(-1, 3), torch.tensor(dx), torch.tensor(bx), torch.tensor([float(v) for v in nx]))
    n = [int(v) for v in nx]
    nvox = n[0] * n[1] * n[2]
    rank = torch.where(kept, (idx[:, 0] * n[1] + idx[:, 1]) * n[2] + idx[:, 2], torch.full_like(idx[:, 0], nvox))
    order = torch.argsort(rank, stable=True)
    counts = torch.bincount(rank[kept], minlength=nvox)
    start = torch.zeros(nvox + 1, dtype=torch.int64)
    start[1:] = torch.cumsum(counts, 0)
    return SplatIndex(order, start, None, n[0], n[1], n[2], 1, geom.reshape(-1, 3).shape[0])


def splat_index_slab(index: SplatIndex, x0: int, x1: int) -> SplatIndex:
    per_x = index.ny * index.nz
    return SplatIndex(index.order, index.voxel_start[x0 * per_x: x1 * per_x + 1], None, x1 - x0, index.ny, index.nz, 1, index.P)


def lift_splat(depth_prob, img_feat, index: SplatIndex, out=None):
    B, D, H, W = depth_prob.shape
    C = img_feat.shape[-1]
    start = index.voxel_start
    nvox = start.numel() - 1
    pts = index.order[int(start[0]): int(start[-1])]
    vox = torch.repeat_interleave(torch.arange(nvox), (start[1:] - start[:-1]))
    pix = pts % (H * W)
    vals = depth_prob.reshape(-1).double()[pts][:, None] * img_feat.reshape(H * W, C).double()[pix]
    res = torch.zeros((nvox, C), dtype=torch.float64).index_add_(0, vox, vals).view(1, index.nx, index.ny, index.nz, C)
    if out is None:
        return res.to(img_feat.dtype)
    out.copy_(res)
    return out


def trilinear(x, size, want_labels=False):
    y = F.interpolate(x.double().permute(0, 4, 1, 2, 3), size=tuple(size), mode="trilinear", align_corners=False)
    ycl = y.permute(0, 2, 3, 4, 1).contiguous().to(x.dtype)
    return ycl, (ycl.argmax(-1).to(torch.uint8) if want_labels else None)


@contextlib.contextmanager
def math_scope(group):
    yield
