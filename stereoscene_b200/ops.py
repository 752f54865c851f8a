"""Tensor-level wrappers over the C ABI (include/stereoscene_b200.h).

PyTorch is used here for device memory, streams and views only; every arithmetic step of the
hot path is a call into libstereoscene_b200.so.  There is no fallback: a CPU tensor, a wrong
dtype or a missing library raises.

Data model
----------
``Vol`` is a channels-last volume ``data[B,D,H,W,C]`` (possibly a channel slice of a wider
buffer) together with a *pending affine*: per-(batch,channel) ``scale``/``shift`` and an
activation that the consumer applies while loading, i.e. the logical value of the volume is
``act(data*scale+shift)``.  GroupNorm, eval-mode BatchNorm, SE gates and the CA3D gate are all
expressed this way, so no normalisation pass over a volume ever runs.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import cabi
from .cabi import SS_ACT_GELU, SS_ACT_NONE, SS_ACT_RELU, SS_ACT_SIGMOID, SS_ACT_SWISH, SS_MATH_3XTF32, SS_MATH_F16, SS_MATH_F16X3, SS_MATH_TF32, SS_MATH_TF32X3  # noqa: F401

_DEFAULT_MATH = SS_MATH_TF32
_USE_TCGEN05 = os.environ.get("STEREOSCENE_B200_NO_TCGEN05", "0") != "1"
_FUSE_JOIN = os.environ.get("STEREOSCENE_B200_NO_FUSED_JOIN", "0") != "1"      # A/B switch for ops.conv_join
_USE_F16X3 = os.environ.get("STEREOSCENE_B200_NO_F16X3", "0") != "1"           # A/B switch: single-launch fp16-split compensation


def use_f16x3(flag: bool):
    """Serve SS_MATH_TF32X3 layers of the halo-resident / box kernels with the single-launch fp16-split variant
    (default) or with three accumulating TF32 launches like every other kernel family."""
    global _USE_F16X3
    _USE_F16X3 = bool(flag)


def use_tcgen05(flag: bool):
    """Route eligible convolutions through the tcgen05 kernel (default) or keep every layer on the
    mma.sync kernel (used by tests to cross-check the two implementations)."""
    global _USE_TCGEN05
    _USE_TCGEN05 = bool(flag)


def set_default_math(mode: int):
    """Uniform math mode for every stage (switches the per-stage policy off): SS_MATH_TF32 (fast), SS_MATH_TF32X3 (error-compensated split TF32 on the tcgen05 kernels, ~fp32 accuracy at ~3x the
    tensor work) or SS_MATH_3XTF32 (the same compensation on the mma.sync kernels) for conv / BRI energy."""
    global _DEFAULT_MATH, _POLICY
    assert mode in (SS_MATH_TF32, SS_MATH_3XTF32, SS_MATH_TF32X3, SS_MATH_F16)
    _DEFAULT_MATH = mode
    _POLICY = None                      # a uniform mode replaces the per-stage policy


def default_math() -> int:
    return _DEFAULT_MATH


# ---- per-stage math policy -----------------------------------------------------------------------------------------
# The path has four stage groups: "stereo" (stereofeature_net, cost volume, cost aggregation), "depthnet", "mie"
# (BRI + DVE) -- the frustum-space stages, each ending in a softmax over depth that amplifies an absolute error of the
# depth logits into a relative error of the probabilities -- and "voxel" (3-D encoder, neck, head).  Measured at
# configs[2] against the reference's own forward (profiles/r02_parity_split_experiment.txt): with depth_net and the MIE
# block compensated every voxel-space stage stays within 1e-3 of the reference (both error norms) in plain TF32;
# compensating the stereo branch or the voxel stages on top changes the logits error by < 10 %, compensating only one
# of depth_net / MIE leaves the logits at 1.4e-3 rms.  A policy maps group -> math mode; the modules enter
# ``math_scope(group)`` around their stage.
MATH_POLICIES = {
    "tf32": {},                                                                          # every stage plain TF32
    # the parity-green product mode (CA3D sits on a residual branch: leaving it in TF32 moves the logits error 6.8e-4 -> 7.4e-4);
    # the voxel stack's halo-resident layers (encoder blocks, head) multiply fp16 operands: TF32's significand, half the MMAs
    # "image" = the 2-D image encoder in front of the path (row N2): ~110 chained pointwise GEMMs -- plain TF32 leaves its output
    # features at 8e-4 of the reference before the volumetric stages add theirs, so it runs compensated
    # (the stereo branch's and CA3D's halo-resident layers take fp16 operands too: same significand as TF32, measured the same
    # logits error -- 4.6e-4 / 7.6e-4 at configs[2] -- and 0.11 ms less; "mixed_tf32stereo" is the policy without that)
    "mixed": {"image": SS_MATH_TF32X3, "stereo": SS_MATH_F16, "depthnet": SS_MATH_TF32X3, "mie": SS_MATH_TF32X3, "mie.ca3d": SS_MATH_F16, "voxel": SS_MATH_F16},
    "mixed_tf32stereo": {"image": SS_MATH_TF32X3, "depthnet": SS_MATH_TF32X3, "mie": SS_MATH_TF32X3, "mie.ca3d": SS_MATH_TF32, "voxel": SS_MATH_F16},
    "mixed_tf32voxel": {"image": SS_MATH_TF32X3, "depthnet": SS_MATH_TF32X3, "mie": SS_MATH_TF32X3, "mie.ca3d": SS_MATH_TF32},
    # alias of "mixed" since the end of round 2 (kept: the parity tests and earlier profiles name it)
    "mixed16": {"image": SS_MATH_TF32X3, "stereo": SS_MATH_F16, "depthnet": SS_MATH_TF32X3, "mie": SS_MATH_TF32X3, "mie.ca3d": SS_MATH_F16, "voxel": SS_MATH_F16},
    "f16": {g: SS_MATH_F16 for g in ("image", "stereo", "depthnet", "mie", "voxel")},
    "tf32x3": {g: SS_MATH_TF32X3 for g in ("image", "stereo", "depthnet", "mie", "voxel")},
    "3xtf32": {g: SS_MATH_3XTF32 for g in ("image", "stereo", "depthnet", "mie", "voxel")},
}
DEFAULT_POLICY = "mixed"
_POLICY: Optional[dict] = dict(MATH_POLICIES[DEFAULT_POLICY])


def set_math_policy(policy):
    """``policy``: a name from MATH_POLICIES, a dict group -> SS_MATH_*, or None = the product default ("mixed":
    the cheapest policy whose logits are within 1e-3 of the reference at configs[1] and configs[2])."""
    global _POLICY, _DEFAULT_MATH
    if policy is None:
        policy = DEFAULT_POLICY
    if isinstance(policy, str):
        policy = MATH_POLICIES[policy]
    _POLICY = dict(policy)
    _DEFAULT_MATH = SS_MATH_TF32


def math_policy():
    return _POLICY


class math_scope:
    """Context manager: run a stage group in the math mode the active policy assigns to it (no policy: unchanged).
    ``group`` may be a sub-stage "parent.child" ("depthnet.depth", "mie.hourglass", ...): a policy entry for the
    sub-stage wins over the entry of its parent."""

    def __init__(self, group: str):
        self.group = group

    def __enter__(self):
        global _DEFAULT_MATH
        self.saved = _DEFAULT_MATH
        if _POLICY is not None:
            parent = self.group.split(".", 1)[0]
            _DEFAULT_MATH = _POLICY.get(self.group, _POLICY.get(parent, SS_MATH_TF32))
        return self

    def __exit__(self, *exc):
        global _DEFAULT_MATH
        _DEFAULT_MATH = self.saved
        return False


# name -> mode constant, for the tests / bench (--math)
MATH_MODES = {"tf32": SS_MATH_TF32, "f16": SS_MATH_F16, "tf32x3": SS_MATH_TF32X3, "3xtf32": SS_MATH_3XTF32}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")


def _vol_ldc(t: torch.Tensor, name: str) -> int:
    """Voxel stride of a [B,D,H,W,C] tensor that is a channel slice of a contiguous buffer."""
    _need_cuda_f32(t, name)
    if t.dim() != 5:
        raise RuntimeError(f"{name}: expected [B,D,H,W,C], got {tuple(t.shape)}")
    B, D, H, W, Cc = t.shape
    sb, sd, sh, sw, sc = t.stride()
    ldc = sw if W > 1 else (sh if H > 1 else (sd if D > 1 else (sb if B > 1 else Cc)))
    ok = (sc == 1 or Cc == 1) and (W == 1 or sw == ldc) and (H == 1 or sh == W * ldc) and \
         (D == 1 or sd == H * W * ldc) and (B == 1 or sb == D * H * W * ldc) and ldc >= Cc
    if not ok:
        raise RuntimeError(f"{name}: not a channels-last volume (shape {tuple(t.shape)}, stride {t.stride()})")
    return int(ldc)


@dataclass
class Vol:
    data: torch.Tensor                      # [B,D,H,W,C]
    scale: Optional[torch.Tensor] = None    # [B,C] contiguous
    shift: Optional[torch.Tensor] = None
    act: int = SS_ACT_NONE

    @property
    def shape(self):
        return self.data.shape

    @property
    def C(self) -> int:
        return self.data.shape[-1]

    @property
    def is_plain(self) -> bool:
        return self.scale is None and self.act == SS_ACT_NONE

    def plain(self) -> torch.Tensor:
        """The logical value as a contiguous [B,D,H,W,C] tensor (materialises if pending)."""
        if self.is_plain:
            return self.data
        return join(self, None, out_act=SS_ACT_NONE)

    def ncdhw(self) -> torch.Tensor:
        """Logical [B,C,D,H,W] view (channels_last_3d memory) -- the reference's tensor layout."""
        return self.plain().permute(0, 4, 1, 2, 3)


class StatsArena:
    """Zero-initialised double[B][C][2] blocks for the per-channel sums, handed out from one pool
    that is cleared with a single memset per forward."""

    def __init__(self, device, capacity=1 << 16):
        self.device = device
        self.capacity = capacity
        self.pool = torch.zeros(capacity, dtype=torch.float64, device=device)
        self.used = 0

    def reset(self):
        if self.used:
            self.pool[: self.used].zero_()
        self.used = 0

    def take(self, B: int, Cc: int) -> torch.Tensor:
        n = B * Cc * 2
        if self.used + n > self.capacity:
            # grow: a fresh zeroed pool (old blocks stay alive through their views)
            self.capacity = max(self.capacity * 2, n * 2)
            self.pool = torch.zeros(self.capacity, dtype=torch.float64, device=self.device)
            self.used = 0
        blk = self.pool[self.used: self.used + n].view(B, Cc, 2)
        self.used += n
        return blk


_arenas = {}


def arena(device) -> StatsArena:
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    a = _arenas.get(key)
    if a is None:
        a = _arenas[key] = StatsArena(device)
    return a


# ------------------------------------------------------------------------------------------
# convolution family
# ------------------------------------------------------------------------------------------
class PackedConv:
    """Geometry + weights of one Conv3d / ConvTranspose3d / Conv2d in the kernel's layout
    (float[taps][Cin][Cout_padded]); built once from the nn.Module that owns the parameter (so the
    state_dict stays the reference's) and refreshed if the parameter is reassigned or modified."""

    def __init__(self, module: torch.nn.Module):
        self.module = module
        m = module
        if isinstance(m, torch.nn.Conv2d):
            self.k = (1,) + tuple(m.kernel_size); self.s = (1,) + tuple(m.stride)
            self.p = (0,) + tuple(m.padding); self.d = (1,) + tuple(m.dilation)
            self.transposed = False; self.outpad = (0, 0, 0)
        elif isinstance(m, torch.nn.ConvTranspose3d):
            self.k = tuple(m.kernel_size); self.s = tuple(m.stride); self.p = tuple(m.padding)
            self.d = tuple(m.dilation); self.transposed = True; self.outpad = tuple(m.output_padding)
        elif isinstance(m, torch.nn.Conv3d):
            self.k = tuple(m.kernel_size); self.s = tuple(m.stride); self.p = tuple(m.padding)
            self.d = tuple(m.dilation); self.transposed = False; self.outpad = (0, 0, 0)
        else:
            raise TypeError(type(m))
        if m.groups != 1:
            raise RuntimeError("grouped convolution is not on the hot path")
        self.Cin, self.Cout = m.in_channels, m.out_channels
        # rows of the packed weight matrices: narrow layers are padded to one 32-column tensor-core tile
        self.CoutP = 32 if self.Cout <= 32 else (self.Cout + 7) // 8 * 8
        self._key = None
        self._w = None
        self._kkey = None
        self._wk = None
        self._skey = None
        self._ws = None
        self._hkey = None
        self._wh = None

    def weights(self) -> torch.Tensor:
        w = self.module.weight
        key = (w.data_ptr(), w._version, w.device)
        if key != self._key:
            with torch.no_grad():
                wd = w.detach()
                if wd.dim() == 4:
                    wd = wd.unsqueeze(2)
                if self.transposed:          # [Cin,Cout,kd,kh,kw] -> [kd,kh,kw,Cin,Cout]
                    pk = wd.permute(2, 3, 4, 0, 1)
                else:                        # [Cout,Cin,kd,kh,kw] -> [kd,kh,kw,Cin,Cout]
                    pk = wd.permute(2, 3, 4, 1, 0)
                pk = pk.reshape(-1, self.Cin, self.Cout).float()
                if self.CoutP != self.Cout:
                    pk = torch.nn.functional.pad(pk, (0, self.CoutP - self.Cout))
                self._w = pk.contiguous()
            self._key = key
        return self._w

    def weights_kmajor(self) -> torch.Tensor:
        """float[taps][Cout_padded][Cin], rounded to TF32 (nearest, ties away) -- tcgen05 B operand."""
        w = self.module.weight
        key = (w.data_ptr(), w._version, w.device)
        if key != self._kkey:
            with torch.no_grad():
                wd = w.detach()
                if wd.dim() == 4:
                    wd = wd.unsqueeze(2)
                pk = wd.permute(2, 3, 4, 1, 0) if self.transposed else wd.permute(2, 3, 4, 0, 1)   # [k,k,k,Cout,Cin]
                pk = pk.reshape(-1, self.Cout, self.Cin).float()
                if self.CoutP != self.Cout:
                    pk = torch.nn.functional.pad(pk, (0, 0, 0, self.CoutP - self.Cout))
                bits = pk.contiguous().view(torch.int32)
                bits = (bits + 0x1000) & ~0x1FFF                      # cvt.rna.tf32.f32 on the magnitude bits
                self._wk = bits.view(torch.float32).contiguous()
            self._kkey = key
        return self._wk

    def weights_kmajor_split(self) -> torch.Tensor:
        """float[2][taps][Cout_padded][Cin]: hi = tf32(w) followed by lo = tf32(w - hi) -- the B operands of the
        compensated tcgen05 mode (SS_MATH_TF32X3)."""
        w = self.module.weight
        key = (w.data_ptr(), w._version, w.device)
        if key != self._skey:
            with torch.no_grad():
                wd = w.detach()
                if wd.dim() == 4:
                    wd = wd.unsqueeze(2)
                pk = wd.permute(2, 3, 4, 1, 0) if self.transposed else wd.permute(2, 3, 4, 0, 1)   # [k,k,k,Cout,Cin]
                pk = pk.reshape(-1, self.Cout, self.Cin).float()
                if self.CoutP != self.Cout:
                    pk = torch.nn.functional.pad(pk, (0, 0, 0, self.CoutP - self.Cout))
                pk = pk.contiguous()

                def rna(t):
                    return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

                hi = rna(pk)
                lo = rna((pk - hi).contiguous())
                self._ws = torch.stack([hi, lo]).contiguous()
            self._skey = key
        return self._ws

    def weights_kmajor_f16(self):
        """(packed, acc_scale) for SS_MATH_F16X3: the K-major array of ``weights_kmajor`` in which every 128-byte row (one
        output channel's 32-channel chunk) holds 64 fp16 values, hi = fp16(w / acc_scale) of the 32 channels followed by
        lo = fp16(w / acc_scale - hi); acc_scale is the power of two that brings max|w| to ~2^10, so that the lo halves
        of all but negligible weights are normal fp16 numbers."""
        w = self.module.weight
        key = (w.data_ptr(), w._version, w.device)
        if key != self._hkey:
            with torch.no_grad():
                wd = w.detach()
                if wd.dim() == 4:
                    wd = wd.unsqueeze(2)
                pk = wd.permute(2, 3, 4, 1, 0) if self.transposed else wd.permute(2, 3, 4, 0, 1)   # [k,k,k,Cout,Cin]
                pk = pk.reshape(-1, self.Cout, self.Cin).float()
                if self.CoutP != self.Cout:
                    pk = torch.nn.functional.pad(pk, (0, 0, 0, self.CoutP - self.Cout))
                amax = float(pk.abs().max())
                e = math.floor(math.log2(amax)) if amax > 0.0 else 0
                acc_scale = 2.0 ** (e - 10)
                ws = pk * (1.0 / acc_scale)
                hi = ws.half()
                lo = (ws - hi.float()).half()
                T, Cp, Ci = ws.shape
                both = torch.cat([hi.view(T, Cp, Ci // 32, 32), lo.view(T, Cp, Ci // 32, 32)], dim=-1).contiguous()
                self._wh = (both.view(T, Cp, Ci * 2).view(torch.float32).contiguous(), acc_scale)
            self._hkey = key
        return self._wh

    def out_size(self, din: Sequence[int], pad: Optional[Sequence[int]] = None) -> Tuple[int, int, int]:
        p = self.p if pad is None else pad
        out = []
        for i in range(3):
            if self.transposed:
                out.append((din[i] - 1) * self.s[i] - 2 * p[i] + self.d[i] * (self.k[i] - 1) + self.outpad[i] + 1)
            else:
                out.append((din[i] + 2 * p[i] - self.d[i] * (self.k[i] - 1) - 1) // self.s[i] + 1)
        return tuple(out)


_packed_cache = {}


def packed(module: torch.nn.Module) -> PackedConv:
    pc = _packed_cache.get(id(module))
    if pc is None or pc.module is not module:
        pc = _packed_cache[id(module)] = PackedConv(module)
    return pc


def _halo_or_march_layer(pc: "PackedConv", Din: int, Hin: int, Win: int, Cin: int) -> bool:
    """Mirror of the dispatch in ss_conv3d_tc_fwd (conv3d_halo.cu:try_conv_halo, conv3d_march.cu:try_conv_march32): does
    this stride-1 3x3(x3) layer run on a kernel that applies a pending affine once per landed plane?  (Only used to
    decide whether materialising a small pending input pays; a mismatch costs time, never correctness.)"""
    if pc.transposed or pc.s != (1, 1, 1) or pc.d != (1, 1, 1) or pc.k[1:] != (3, 3) or pc.p[1:] != (1, 1):
        return False
    if not ((pc.k[0] == 3 and pc.p[0] == 1) or (pc.k[0] == 1 and pc.p[0] == 0)):
        return False
    if Cin == 32 and pc.CoutP == 32 and pc.k[0] == 3:
        return Win >= 8 and Hin >= 8 and Din >= 3
    if pc.CoutP < 64:
        return False
    tiles = lambda h, w: ((h + 31) // 32) * ((w + 7) // 8)        # noqa: E731
    return Hin * Win / (256.0 * min(tiles(Hin, Win), tiles(Win, Hin))) >= 0.7


def _halo_layer(pc: "PackedConv", Din: int, Hin: int, Win: int, Cin: int) -> bool:
    """Does this layer run on the halo-resident kernel (conv3d_halo.cu:try_conv_halo)?  The fp16 single pass (SS_MATH_F16) is
    only a gain there (weight tiles of half the size in a deeper ring, a third plane slot): every other kernel keeps TF32,
    which has the same significand."""
    if Cin == 32 and pc.CoutP == 32:
        return False
    return pc.CoutP >= 64 and _halo_or_march_layer(pc, Din, Hin, Win, Cin)


_MATERIALIZE_BYTES = 16 << 20


def conv(x: Vol, module: torch.nn.Module, out: Optional[torch.Tensor] = None, out_act: int = SS_ACT_NONE,
         want_stats: bool = False, math_mode: Optional[int] = None, use_bias: bool = True,
         pad: Optional[Sequence[int]] = None, stats_planes: Optional[Tuple[int, int]] = None, accumulate: bool = False,
         splitk_ws: Optional[torch.Tensor] = None):
    """y = act_out(conv(act_in(x*scale+shift)) + bias); returns (y [B,D',H',W',Cout], stats or None).
    ``accumulate``: y = act_out(conv + bias + out) in place on ``out`` (the identity shortcut of a residual block, taken in the
    epilogue of the tcgen05 kernels; on the mma.sync kernels it is a separate join).  ``splitk_ws``: a float32 scratch tensor the
    tcgen05 box kernel may use to split a long K over several CTAs per tile (small-M layers of the image encoder).
    ``out`` may be a channel slice of a concatenation buffer.  ``pad`` overrides the module's padding and
    ``stats_planes`` = (d0, d1) restricts the GroupNorm sums to output planes d0 <= d < d1: both serve the X-slab
    sharded mode (stereoscene_b200.xshard), where a rank convolves its slab plus halo planes."""
    lib = cabi.load()
    pc = packed(module)
    xin = x.data
    in_ldc = _vol_ldc(xin, "conv input")
    B, Din, Hin, Win, Cin = xin.shape
    if Cin != pc.Cin:
        raise RuntimeError(f"conv: input has {Cin} channels, layer expects {pc.Cin}")
    pad_eff = tuple(pc.p) if pad is None else tuple(int(v) for v in pad)
    Do, Ho, Wo = pc.out_size((Din, Hin, Win), pad_eff)
    sd0, sd1 = (0, 0) if stats_planes is None else (int(stats_planes[0]), int(stats_planes[1]))
    if out is None:
        out = torch.empty((B, Do, Ho, Wo, pc.Cout), dtype=torch.float32, device=xin.device)
    elif tuple(out.shape) != (B, Do, Ho, Wo, pc.Cout):
        raise RuntimeError(f"conv: out has shape {tuple(out.shape)}, expected {(B, Do, Ho, Wo, pc.Cout)}")
    out_ldc = _vol_ldc(out, "conv output")
    stats = arena(xin.device).take(B, pc.Cout) if want_stats else None
    bias = module.bias if (use_bias and module.bias is not None) else None
    if x.scale is not None:
        if tuple(x.scale.shape) != (B, Cin) or not x.scale.is_contiguous() or not x.shift.is_contiguous():
            raise RuntimeError("conv: pending affine must be contiguous [B,Cin]")
    mm = _DEFAULT_MATH if math_mode is None else math_mode
    d = cabi.ConvDesc(B, Din, Hin, Win, Cin, Do, Ho, Wo, pc.Cout, *pc.k, *pc.s, *pad_eff, *pc.d,
                      1 if pc.transposed else 0, in_ldc, out_ldc, x.act, out_act, mm, pc.CoutP, sd0, sd1)
    # single-output-channel layers ride the tensor-core kernel too (N padded to 32): it is faster than the FMA kernel
    tc = (_USE_TCGEN05 and mm in (SS_MATH_TF32, SS_MATH_TF32X3, SS_MATH_F16) and Cin % 32 == 0 and ((pc.Cout % 4 == 0 and pc.Cout >= 32) or pc.Cout < 32)
          and in_ldc % 4 == 0 and xin.data_ptr() % 16 == 0)
    d.acc_scale = 1.0
    if accumulate:
        if out is None or want_stats:
            raise RuntimeError("conv: accumulate needs an explicit out tensor and is not combined with statistics")
        if not tc or mm == SS_MATH_3XTF32:     # mma.sync kernels: convolve into a temporary, then one join
            tmp, _ = conv(x, module, None, SS_ACT_NONE, False, math_mode, use_bias, pad)
            join(Vol(tmp), Vol(out), out_act=out_act, out=out)
            return out, None
        d.accumulate = 1
    if splitk_ws is not None and tc:
        d.splitk_ws, d.splitk_ws_bytes = splitk_ws.data_ptr(), splitk_ws.numel() * 4
    if mm == SS_MATH_TF32X3 and not tc:        # layers the tcgen05 kernels do not take (Cin = 2, odd strides): mma.sync split TF32
        mm = SS_MATH_3XTF32
        d.math = mm
    if (tc and not x.is_plain and math.prod(pc.k) >= 27 and Cin >= 256 and xin.numel() * 4 <= _MATERIALIZE_BYTES
            and not _halo_or_march_layer(pc, Din, Hin, Win, Cin)):
        # the per-tap box kernel re-applies a pending affine for every tap (27 x the work of the plane kernels): for a
        # small volume one elementwise pass first is cheaper (512-channel 32x32x4 layers: 0.18 -> 0.11 ms + 0.005)
        x = Vol(x.plain())
        xin = x.data
        in_ldc = _vol_ldc(xin, "conv input")
        d = cabi.ConvDesc(B, Din, Hin, Win, Cin, Do, Ho, Wo, pc.Cout, *pc.k, *pc.s, *pad_eff, *pc.d,
                          1 if pc.transposed else 0, in_ldc, out_ldc, x.act, out_act, mm, pc.CoutP, sd0, sd1)
        d.acc_scale = 1.0
        d.accumulate = 1 if accumulate else 0
        if splitk_ws is not None:
            d.splitk_ws, d.splitk_ws_bytes = splitk_ws.data_ptr(), splitk_ws.numel() * 4
    if mm == SS_MATH_F16 and not (tc and pad is None and _halo_layer(pc, Din, Hin, Win, Cin)):
        mm = SS_MATH_TF32            # only the halo-resident kernel gains from fp16 operands; TF32 has the same 11-bit significand
        d.math = mm
    if tc and ((mm == SS_MATH_TF32X3 and _USE_F16X3 and lib.ss_conv3d_tc_f16x3_supported(C.byref(d)) == 1) or mm == SS_MATH_F16):
        # the halo-resident / box kernels take fp16 operands in ONE launch: the compensated hi/lo split (1.5x the TF32 tensor
        # work) or the hi halves only (SS_MATH_F16: half the TF32 tensor work, same significand)
        wk, d.acc_scale = pc.weights_kmajor_f16()
        d.math = SS_MATH_F16 if mm == SS_MATH_F16 else SS_MATH_F16X3
        rc = lib.ss_conv3d_tc_fwd(C.byref(d), xin.data_ptr(), _ptr(x.scale), _ptr(x.shift), wk.data_ptr(),
                                  _ptr(bias.detach() if bias is not None else None), out.data_ptr(), _ptr(stats), _stream())
        cabi.check(rc, "ss_conv3d_tc_fwd")
        return out, stats
    if tc:
        wk = pc.weights_kmajor_split() if mm == SS_MATH_TF32X3 else pc.weights_kmajor()
        rc = lib.ss_conv3d_tc_fwd(C.byref(d), xin.data_ptr(), _ptr(x.scale), _ptr(x.shift), wk.data_ptr(),
                                  _ptr(bias.detach() if bias is not None else None), out.data_ptr(), _ptr(stats), _stream())
        cabi.check(rc, "ss_conv3d_tc_fwd")
    else:
        rc = lib.ss_conv3d_fwd(C.byref(d), xin.data_ptr(), _ptr(x.scale), _ptr(x.shift), pc.weights().data_ptr(),
                               _ptr(bias.detach() if bias is not None else None), out.data_ptr(), _ptr(stats), _stream())
        cabi.check(rc, "ss_conv3d_fwd")
    return out, stats


def conv_join(x: Vol, module: torch.nn.Module, out_affine: Optional[Vol], res: Optional[Vol], out_act: int = SS_ACT_NONE,
              use_bias: bool = True) -> torch.Tensor:
    """out_act( A_out(conv(x)) + A_res(res) ) as ONE kernel where the layer qualifies (stride-2 transposed k3 convs: the
    hourglass up-convolutions), else conv followed by the join kernel.  ``out_affine`` carries only scale / shift / act of
    the convolution result (e.g. ``bn_pending(None-like)``): its ``data`` is ignored.  Returns the plain joined tensor."""
    lib = cabi.load()
    pc = packed(module)
    xin = x.data
    B, Din, Hin, Win, Cin = xin.shape
    Do, Ho, Wo = pc.out_size((Din, Hin, Win))
    mm = SS_MATH_TF32 if _DEFAULT_MATH == SS_MATH_F16 else _DEFAULT_MATH      # the transposed kernel has no fp16 operand path
    fused_ok = (_USE_TCGEN05 and _FUSE_JOIN and mm in (SS_MATH_TF32, SS_MATH_TF32X3) and Cin % 32 == 0 and xin.data_ptr() % 16 == 0 and
                (out_affine is None or out_affine.act == SS_ACT_NONE))
    if fused_ok:
        in_ldc = _vol_ldc(xin, "conv_join input")
        out = torch.empty((B, Do, Ho, Wo, pc.Cout), dtype=torch.float32, device=xin.device)
        d = cabi.ConvDesc(B, Din, Hin, Win, Cin, Do, Ho, Wo, pc.Cout, *pc.k, *pc.s, *pc.p, *pc.d, 1 if pc.transposed else 0,
                          in_ldc, pc.Cout, x.act, out_act, mm, pc.CoutP)
        fused_ok = in_ldc % 4 == 0 and bool(lib.ss_conv3d_tc_join_supported(C.byref(d)))
    if not fused_ok:
        y, _ = conv(x, module, use_bias=use_bias)
        yv = Vol(y) if out_affine is None else Vol(y, out_affine.scale, out_affine.shift, out_affine.act)
        return join(yv, res, out_act=out_act)
    if res is not None and tuple(res.data.shape) != tuple(out.shape):
        raise RuntimeError(f"conv_join: residual has shape {tuple(res.data.shape)}, expected {tuple(out.shape)}")
    for v in (out_affine, res):
        if v is not None and v.scale is not None and (tuple(v.scale.shape) != (B, pc.Cout) or not v.scale.is_contiguous()
                                                      or not v.shift.is_contiguous()):
            raise RuntimeError("conv_join: affines must be contiguous [B,Cout]")
    j = cabi.ConvJoin(_ptr(out_affine.scale) if out_affine is not None else None,
                      _ptr(out_affine.shift) if out_affine is not None else None,
                      _ptr(res.data) if res is not None else None, _ptr(res.scale) if res is not None else None,
                      _ptr(res.shift) if res is not None else None,
                      _vol_ldc(res.data, "conv_join residual") if res is not None else 0, res.act if res is not None else 0)
    bias = module.bias if (use_bias and module.bias is not None) else None
    d.acc_scale = 1.0
    if mm == SS_MATH_TF32X3 and _USE_F16X3:
        wk, d.acc_scale = pc.weights_kmajor_f16()
        d.math = SS_MATH_F16X3
    else:
        wk = pc.weights_kmajor_split() if mm == SS_MATH_TF32X3 else pc.weights_kmajor()
    rc = lib.ss_conv3d_tc_join_fwd(C.byref(d), xin.data_ptr(), _ptr(x.scale), _ptr(x.shift), wk.data_ptr(),
                                   _ptr(bias.detach() if bias is not None else None), C.byref(j), out.data_ptr(), _stream())
    cabi.check(rc, "ss_conv3d_tc_join_fwd")
    return out


def voxels_per_channel(t: torch.Tensor) -> int:
    return int(t.shape[1] * t.shape[2] * t.shape[3])


def gn_pending(y: torch.Tensor, stats: torch.Tensor, gn: torch.nn.GroupNorm, act: int = SS_ACT_NONE,
               scale_out: Optional[torch.Tensor] = None, shift_out: Optional[torch.Tensor] = None,
               count: Optional[float] = None) -> Vol:
    """Turn a producer's sums into the pending GroupNorm(+activation) of its raw output.  ``count`` = voxels per
    (batch, channel) the sums cover (default: all of ``y``; the X-slab sharded mode passes the global count after
    all-reducing the sums of the slabs)."""
    lib = cabi.load()
    B, Cc = stats.shape[0], stats.shape[1]
    if scale_out is None:
        ss = torch.empty((2, B, Cc), dtype=torch.float32, device=y.device)
        scale_out, shift_out = ss[0], ss[1]
    ld = scale_out.stride(0) if B > 1 else max(Cc, scale_out.stride(0))
    rc = lib.ss_gn_finalize(stats.data_ptr(), gn.weight.data_ptr(), gn.bias.data_ptr(), B, Cc, gn.num_groups,
                            float(voxels_per_channel(y) if count is None else count), float(gn.eps), scale_out.data_ptr(), shift_out.data_ptr(),
                            int(ld), _stream())
    cabi.check(rc, "ss_gn_finalize")
    return Vol(y, scale_out, shift_out, act)


_bn_cache = {}


def bn_pending(y: torch.Tensor, bn: torch.nn.modules.batchnorm._BatchNorm, act: int = SS_ACT_NONE) -> Vol:
    """Eval-mode BatchNorm as a pending affine from the running statistics (host-side, cached:
    it depends on parameters only)."""
    B = y.shape[0]
    key = (bn.weight.data_ptr(), bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, bn.weight.device)
    slot = (id(bn), B)                                  # one entry per (module, batch size): alternating batch sizes do not thrash
    hit = _bn_cache.get(slot)
    if hit is None or hit[0] != key or hit[3] is not bn:
        with torch.no_grad():
            sc = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
            sh = bn.bias.detach().float() - bn.running_mean.float() * sc
            hit = (key, sc.unsqueeze(0).repeat(B, 1).contiguous(), sh.unsqueeze(0).repeat(B, 1).contiguous(), bn)
        _bn_cache[slot] = hit
    return Vol(y, hit[1], hit[2], act)


def gn_pending_gated(y: torch.Tensor, stats: torch.Tensor, gn: torch.nn.GroupNorm, act: int, gate1: torch.Tensor,
                     gate2: Optional[torch.Tensor] = None):
    """Pending GroupNorm(+activation) of a raw output, multiplied by one or two positive per-(batch,channel) gates:
    returns one Vol per gate (same data, scale*gate_k / shift*gate_k) from ONE launch."""
    lib = cabi.load()
    B, Cc = stats.shape[0], stats.shape[1]
    for g in (gate1, gate2):
        if g is not None and (tuple(g.shape) != (B, Cc) or not g.is_contiguous()):
            raise RuntimeError("gn_pending_gated: gates must be contiguous [B,C]")
    ss = torch.empty((4 if gate2 is not None else 2, B, Cc), dtype=torch.float32, device=y.device)
    rc = lib.ss_gn_finalize_gated(stats.data_ptr(), gn.weight.data_ptr(), gn.bias.data_ptr(), B, Cc, gn.num_groups,
                                  float(voxels_per_channel(y)), float(gn.eps), gate1.data_ptr(), ss[0].data_ptr(), ss[1].data_ptr(),
                                  _ptr(gate2), _ptr(ss[2] if gate2 is not None else None), _ptr(ss[3] if gate2 is not None else None),
                                  _stream())
    cabi.check(rc, "ss_gn_finalize_gated")
    v1 = Vol(y, ss[0], ss[1], act)
    return (v1, Vol(y, ss[2], ss[3], act)) if gate2 is not None else v1


def aspp_pool_shift(stats: torch.Tensor, count: int, w1: torch.Tensor, gn: torch.nn.GroupNorm, w_pool: torch.Tensor,
                    bn_scale: torch.Tensor, bn_shift: torch.Tensor) -> torch.Tensor:
    """shift[B,mid] = bn_shift + bn_scale * (w_pool @ relu(GroupNorm(w1 @ mean_x))) from the per-channel sums of x (one launch)."""
    lib = cabi.load()
    B, Cc = stats.shape[0], stats.shape[1]
    mid = w_pool.shape[0]
    out = torch.empty((2, B, mid), dtype=torch.float32, device=stats.device)
    rc = lib.ss_aspp_pool_shift(stats.data_ptr(), float(count), w1.data_ptr(), gn.weight.data_ptr(), gn.bias.data_ptr(), gn.num_groups,
                                float(gn.eps), w_pool.data_ptr(), bn_scale.data_ptr(), bn_shift.data_ptr(), out[0].data_ptr(),
                                out[1].data_ptr(), B, Cc, mid, _stream())
    out = out[0]
    cabi.check(rc, "ss_aspp_pool_shift")
    return out


_const_cache = {}
_CONST_CACHE_ENTRIES = 64


def cached_const(tag: str, keys: Sequence[torch.Tensor], fn):
    """Value of ``fn()`` cached under the identity (storage address, version, shape) of the tensors it depends on --
    calibration tensors and parameters, which are constant per sequence / checkpoint -- so the handful of tiny host-side ops
    behind gates, packed scalars and the like leaves the steady-state step.  The entry holds the key tensors, so a recycled
    allocation cannot alias the key."""
    key = (tag,) + tuple((t.data_ptr(), t._version, tuple(t.shape), t.device) for t in keys)
    hit = _const_cache.get(key)
    if hit is None:
        with torch.no_grad():
            hit = (fn(), tuple(keys))
        _const_cache[key] = hit
        while len(_const_cache) > _CONST_CACHE_ENTRIES:
            _const_cache.pop(next(iter(_const_cache)))
    return hit[0]


def ca3d_gate(v: Vol, stats: torch.Tensor, conv_reduce: torch.nn.Conv3d, conv_expand: torch.nn.Conv3d) -> Vol:
    """Fold sigmoid(GELU(expand(GELU(reduce(avgpool(v)))))) into v's pending affine (in place)."""
    lib = cabi.load()
    B, Cc = v.scale.shape
    rc = lib.ss_ca3d_gate(stats.data_ptr(), float(voxels_per_channel(v.data)), v.scale.data_ptr(), v.shift.data_ptr(),
                          conv_reduce.weight.data_ptr(), conv_reduce.bias.data_ptr(), conv_expand.weight.data_ptr(),
                          conv_expand.bias.data_ptr(), B, Cc, conv_reduce.out_channels, _stream())
    cabi.check(rc, "ss_ca3d_gate")
    return v


def join(x: Vol, r: Optional[Vol], out_act: int = SS_ACT_NONE, alpha: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act(alpha * x + r) on logical values (x, r pending volumes); plain [B,D,H,W,C] result."""
    lib = cabi.load()
    xd = x.data
    x_ldc = _vol_ldc(xd, "join x")
    B, D, H, W, Cc = xd.shape
    r_ldc = 0
    if r is not None:
        if tuple(r.data.shape) != tuple(xd.shape):
            raise RuntimeError(f"join: shapes differ {tuple(xd.shape)} vs {tuple(r.data.shape)}")
        r_ldc = _vol_ldc(r.data, "join r")
    if out is None:
        out = torch.empty((B, D, H, W, Cc), dtype=torch.float32, device=xd.device)
    out_ldc = _vol_ldc(out, "join out")
    rc = lib.ss_affine_join_fwd(xd.data_ptr(), _ptr(x.scale), _ptr(x.shift), x.act,
                                _ptr(r.data if r is not None else None), _ptr(r.scale if r is not None else None),
                                _ptr(r.shift if r is not None else None), r.act if r is not None else 0,
                                _ptr(alpha.detach() if alpha is not None else None), out_act, B, D * H * W, Cc,
                                x_ldc, r_ldc, out_ldc, out.data_ptr(), _stream())
    cabi.check(rc, "ss_affine_join_fwd")
    return out


def channel_sums(x: Vol) -> torch.Tensor:
    """double[B,C,2] = (sum, sum of squares) over the voxels of the logical value of a pending volume."""
    lib = cabi.load()
    xd = x.data
    ldc = _vol_ldc(xd, "channel_sums")
    B, Cc = xd.shape[0], xd.shape[-1]
    st = arena(xd.device).take(B, Cc)
    rc = lib.ss_channel_sums_fwd(xd.data_ptr(), _ptr(x.scale), _ptr(x.shift), x.act, B, voxels_per_channel(xd), Cc, ldc,
                                 st.data_ptr(), _stream())
    cabi.check(rc, "ss_channel_sums_fwd")
    return st


# ------------------------------------------------------------------------------------------
# 2-D image encoder (row N2): the CUDA-core kernels next to the pointwise GEMMs that ``conv`` serves
# ------------------------------------------------------------------------------------------
def stem_conv2d(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, k: int, stride: int, out_act: int = SS_ACT_NONE) -> torch.Tensor:
    """img [N,Cin<=4,H,W] (the reference's layout) -> channels-last [N,1,ceil(H/s),ceil(W/s),Cout]; TF "SAME" padding;
    w: [k*k*Cin, Cout] ordered (ky, kx, ci), bias [Cout] (BatchNorm folded by the caller)."""
    lib = cabi.load()
    _need_cuda_f32(img, "stem_conv2d")
    img = img.contiguous()
    N, Cin, H, W = img.shape
    Cout = w.shape[1]
    Ho, Wo = -(-H // stride), -(-W // stride)
    y = torch.empty((N, 1, Ho, Wo, Cout), dtype=torch.float32, device=img.device)
    rc = lib.ss_stem_conv2d_fwd(img.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), N, Cin, H, W, Cout, k, stride,
                                out_act, _stream())
    cabi.check(rc, "ss_stem_conv2d_fwd")
    return y


def dwconv2d(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, k: int, stride: int, out_act: int = SS_ACT_NONE,
             want_pool: bool = False, out: Optional[torch.Tensor] = None):
    """Depthwise k x k conv (TF "SAME" padding) of a channels-last [N,1,H,W,C] image; w [k*k, C], bias [C].  Returns
    (y [N,1,ceil(H/s),ceil(W/s),C], pool) with pool = double[N,C,2] whose slot 0 holds the per-image channel sums of y."""
    lib = cabi.load()
    in_ldc = _vol_ldc(x, "dwconv2d input")
    N, D, H, W, Cc = x.shape
    if D != 1:
        raise RuntimeError("dwconv2d: expected a depth-1 volume [N,1,H,W,C]")
    Ho, Wo = -(-H // stride), -(-W // stride)
    if out is None:
        out = torch.empty((N, 1, Ho, Wo, Cc), dtype=torch.float32, device=x.device)
    out_ldc = _vol_ldc(out, "dwconv2d output")
    pool = arena(x.device).take(N, Cc) if want_pool else None
    rc = lib.ss_dwconv2d_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(pool), N, H, W, Cc, in_ldc, out_ldc,
                             k, stride, out_act, _stream())
    cabi.check(rc, "ss_dwconv2d_fwd")
    return out, pool


def se_fc(inp: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int, in_mul: float = 1.0) -> torch.Tensor:
    """out[n,o] = act(bias[o] + in_mul * sum_c in[n,c] w[o,c]); ``inp`` is float[N,Cin] or a double[N,Cin,2] block of channel
    sums (slot 0 is read: with in_mul = 1/pixels that is the squeeze-excite block's pooled mean)."""
    lib = cabi.load()
    stats = inp.dtype == torch.float64
    N, Cin = inp.shape[0], inp.shape[1]
    Cout = w.shape[0]
    if w.shape[1] != Cin or not w.is_contiguous() or not inp.is_contiguous():
        raise RuntimeError(f"se_fc: weight {tuple(w.shape)} does not match input {tuple(inp.shape)}")
    out = torch.empty((N, Cout), dtype=torch.float32, device=inp.device)
    rc = lib.ss_se_fc_fwd(inp.data_ptr(), 1 if stats else 0, w.data_ptr(), _ptr(bias), out.data_ptr(), N, Cin, Cout, float(in_mul),
                          act, _stream())
    cabi.check(rc, "ss_se_fc_fwd")
    return out


def softmax_d(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Softmax over dim 1 of x[B,D,...pixels...] whose per-sample block [D,P] is contiguous
    (a channel slice x[:, :D] of a wider NCHW tensor is fine)."""
    lib = cabi.load()
    _need_cuda_f32(x, "softmax_d")
    B, D = x.shape[0], x.shape[1]
    P = int(math.prod(x.shape[2:]))
    if not x[0].is_contiguous():
        raise RuntimeError("softmax_d: per-sample [D,P] block must be contiguous")
    if out is None:
        out = torch.empty((B, D) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    rc = lib.ss_softmax_d_fwd(x.data_ptr(), x.stride(0) if B > 1 else D * P, out.data_ptr(),
                              out.stride(0) if B > 1 else D * P, B, D, P, _stream())
    cabi.check(rc, "ss_softmax_d_fwd")
    return out


# ------------------------------------------------------------------------------------------
# stereo cost volume
# ------------------------------------------------------------------------------------------
def disparity_taps(calib: torch.Tensor, n_bins: int, down: int = 1):
    """Host-side mirror of the reference's sampling-coordinate arithmetic (warp,
    ViewTransformerLSSVoxel.py:139-150, then grid_sample's align_corners un-normalisation), kept
    in fp32 torch ops so the taps are the reference's to the last bit.  Returns
    (i0 int32 [B,K], w0 float [B,K], w1 float [B,K])."""
    B = calib.shape[0]
    D = n_bins
    k = torch.arange(1, 1 + n_bins // down, dtype=torch.float32, device=calib.device)
    xx = (calib.reshape(B, -1)[:, :1].float() / (down * 4.0)) / k[None, :]
    xn = 2.0 * xx / max(D - 1, 1) - 1.0
    pos = ((xn + 1.0) / 2.0) * (D - 1)
    f = torch.floor(pos)
    w1 = pos - f
    w0 = (f + 1.0) - pos
    i0 = f.clamp(-2.0, float(D + 1)).to(torch.int32)
    return i0.contiguous(), w0.contiguous(), w1.contiguous()


_taps_cache = {}
_TAPS_CACHE_ENTRIES = 8


def _cached_taps(calib: torch.Tensor, n_bins: int):
    """disparity_taps keyed on the calibration tensor (calibration-only, like the splat index: constant
    per sequence, so the handful of tiny host-side ops leaves the steady-state step)."""
    key = (calib.data_ptr(), calib._version, tuple(calib.shape), calib.device, n_bins)
    hit = _taps_cache.get(key)
    if hit is None:
        # the entry keeps `calib` alive, so its storage cannot be recycled for another tensor that would alias the key;
        # an in-place update bumps _version.  Small LRU: sequences / models alternating in one process do not thrash.
        hit = (key, disparity_taps(calib, n_bins), calib)
        _taps_cache[key] = hit
        while len(_taps_cache) > _TAPS_CACHE_ENTRIES:
            _taps_cache.pop(next(iter(_taps_cache)))
    else:
        _taps_cache[key] = _taps_cache.pop(key)          # most recently used last
    return hit[1]


_taps_dev_cache = {}


def _taps_on(device, calib, n_bins, taps):
    """Device copy of host-side taps, cached under the same calibration identity."""
    key = (calib.data_ptr(), calib._version, tuple(calib.shape), calib.device, n_bins, device)
    hit = _taps_dev_cache.get(key)
    if hit is None:
        hit = (tuple(t.to(device) for t in taps), calib)
        _taps_dev_cache[key] = hit
        while len(_taps_dev_cache) > _TAPS_CACHE_ENTRIES:
            _taps_dev_cache.pop(next(iter(_taps_dev_cache)))
    return hit[0]


def cached_state():
    """Every tensor the op wrappers keep in their caches right now (packed weights, BatchNorm affines, disparity taps).
    A CUDA graph bakes their device addresses in, so whoever captures one must hold these references for the life of
    the graph: a cache eviction then cannot hand the memory to another tensor under the graph's feet
    (stereoscene_b200.runtime.VolumetricEngine does)."""
    keep = []
    for pc in _packed_cache.values():
        keep += [pc._w, pc._wk, pc._ws, pc._wh[0] if pc._wh is not None else None]
    for hit in _bn_cache.values():
        keep += [hit[1], hit[2]]
    for hit in list(_taps_cache.values()) + list(_taps_dev_cache.values()):
        keep.append(hit[1] if isinstance(hit[1], tuple) else hit[0])
    keep += [hit[0] for hit in _const_cache.values()]
    return [t for t in keep if t is not None]


def gwc_warp(fea: torch.Tensor, calib: torch.Tensor, maxdisp: int, groups: int) -> torch.Tensor:
    """fea [2B,1,H,W,C] channels-last stereo features (left then right) -> cost volume
    [B,K=maxdisp,H,W,G]."""
    lib = cabi.load()
    _need_cuda_f32(fea, "gwc_warp")
    if not fea.is_contiguous():
        raise RuntimeError("gwc_warp: features must be contiguous channels-last")
    B2, _, H, W, Cc = fea.shape
    B = B2 // 2
    i0, w0, w1 = _cached_taps(calib, maxdisp)            # keyed on the caller's tensor, wherever it lives
    if i0.device != fea.device:
        i0, w0, w1 = _taps_on(fea.device, calib, maxdisp, (i0, w0, w1))
    out = torch.empty((B, maxdisp, H, W, groups), dtype=torch.float32, device=fea.device)
    rc = lib.ss_gwc_warp_fwd(fea.data_ptr(), i0.data_ptr(), w0.data_ptr(), w1.data_ptr(), out.data_ptr(),
                             B, Cc, groups, H, W, maxdisp, maxdisp, _stream())
    cabi.check(rc, "ss_gwc_warp_fwd")
    return out


# ------------------------------------------------------------------------------------------
# BRI attention
# ------------------------------------------------------------------------------------------
def bri_attention(q: torch.Tensor, kv: torch.Tensor, params: torch.Tensor, out: torch.Tensor, out_ld: int,
                  math_mode: Optional[int] = None):
    """q, kv: contiguous [B,D,H,W]; params: device float[7] (wq,bq,wk,bk,wv,bv,gamma);
    out: buffer written at out[b,d,n*out_ld]."""
    lib = cabi.load()
    _need_cuda_f32(q, "bri q"); _need_cuda_f32(kv, "bri kv")
    if not (q.is_contiguous() and kv.is_contiguous()):
        raise RuntimeError("bri_attention: q / kv must be contiguous [B,D,H,W]")
    B, D = q.shape[0], q.shape[1]
    N = int(math.prod(q.shape[2:]))
    wsb = int(lib.ss_bri_workspace_bytes(B, D, N))
    ws = torch.empty(wsb // 4, dtype=torch.float32, device=q.device)
    mm = _DEFAULT_MATH if math_mode is None else math_mode
    if mm in (SS_MATH_TF32X3, SS_MATH_F16):          # BRI on tcgen05 is within 2e-5 of an fp64 evaluation in plain TF32 (softmax algebra in fp32)
        mm = SS_MATH_TF32
    rc = lib.ss_bri_attn_fwd(q.data_ptr(), kv.data_ptr(), params.data_ptr(), ws.data_ptr(), wsb, out.data_ptr(), out_ld,
                             B, D, N, mm, _stream())
    cabi.check(rc, "ss_bri_attn_fwd")
    return out


# ------------------------------------------------------------------------------------------
# lift + splat
# ------------------------------------------------------------------------------------------
@dataclass
class SplatIndex:
    order: torch.Tensor          # int32 [B*P]
    voxel_start: torch.Tensor    # int32 [B*nx*ny*nz + 1]
    coords: Optional[torch.Tensor]   # int32 [B*P,4] (ix,iy,iz,kept)
    nx: int
    ny: int
    nz: int
    B: int
    P: int


def splat_build_index(geom: torch.Tensor, dx, bx, nx: Sequence[int], want_coords: bool = False) -> SplatIndex:
    """geom [B,...,3] ego-frame frustum points -> sorted point index (calibration-only)."""
    lib = cabi.load()
    _need_cuda_f32(geom, "splat geom")
    B = geom.shape[0]
    g = geom.reshape(B, -1, 3).contiguous()
    P = g.shape[1]
    n = [int(round(float(v))) for v in nx]
    dxh = (C.c_float * 3)(*[float(v) for v in dx])
    bxh = (C.c_float * 3)(*[float(v) for v in bx])
    total = B * P
    nvox = B * n[0] * n[1] * n[2]
    order = torch.empty(total, dtype=torch.int32, device=geom.device)
    start = torch.empty(nvox + 1, dtype=torch.int32, device=geom.device)
    coords = torch.empty((total, 4), dtype=torch.int32, device=geom.device) if want_coords else None
    wsb = int(lib.ss_splat_index_workspace_bytes(total))
    ws = torch.empty(wsb, dtype=torch.uint8, device=geom.device)
    rc = lib.ss_splat_build_index(g.data_ptr(), dxh, bxh, n[0], n[1], n[2], B, P, _ptr(coords), order.data_ptr(),
                                  start.data_ptr(), ws.data_ptr(), wsb, _stream())
    cabi.check(rc, "ss_splat_build_index")
    return SplatIndex(order, start, coords, n[0], n[1], n[2], B, P)


def lift_splat(depth_prob: torch.Tensor, img_feat: torch.Tensor, index: SplatIndex, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """depth_prob [B,D,H,W], img_feat [B,H,W,C] (channels-last) -> bev [B,X,Y,Z,C] (``out``: contiguous destination)."""
    lib = cabi.load()
    _need_cuda_f32(depth_prob, "lift depth_prob"); _need_cuda_f32(img_feat, "lift img_feat")
    if not (depth_prob.is_contiguous() and img_feat.is_contiguous()):
        raise RuntimeError("lift_splat: inputs must be contiguous")
    B, D, H, W = depth_prob.shape
    Cc = img_feat.shape[-1]
    if index.B != B or index.P != D * H * W:
        raise RuntimeError("lift_splat: index was built for a different frustum")
    if out is None:
        out = torch.empty((B, index.nx, index.ny, index.nz, Cc), dtype=torch.float32, device=depth_prob.device)
    elif tuple(out.shape) != (B, index.nx, index.ny, index.nz, Cc) or not out.is_contiguous():
        raise RuntimeError("lift_splat: out must be a contiguous [B,X,Y,Z,C] tensor")
    rc = lib.ss_lift_splat_fwd(depth_prob.data_ptr(), img_feat.data_ptr(), index.order.data_ptr(),
                               index.voxel_start.data_ptr(), out.data_ptr(), B, D, H, W, Cc, index.nx, index.ny,
                               index.nz, _stream())
    cabi.check(rc, "ss_lift_splat_fwd")
    return out


def splat_index_slab(index: SplatIndex, x0: int, x1: int) -> SplatIndex:
    """The part of a (single-sample) splat index that fills voxels x0 <= x < x1: the index is sorted by voxel rank
    (x slowest), so an X-slab is a contiguous range of the CSR offsets -- no re-sort, no communication."""
    if index.B != 1:
        raise RuntimeError("splat_index_slab: one sample per index (the sharded mode processes samples one by one)")
    per_x = index.ny * index.nz
    return SplatIndex(index.order, index.voxel_start[x0 * per_x: x1 * per_x + 1], None, x1 - x0, index.ny, index.nz, 1, index.P)


def bev_pool(feats: torch.Tensor, coords: torch.Tensor, B, D, H, W) -> torch.Tensor:
    """Drop-in for ``mmdet3d.ops.bev_pool.bev_pool`` (same argument meaning; B/D/H/W may be 0-d
    tensors as at the reference call site): returns [B,C,D,H,W]."""
    lib = cabi.load()
    B, D, H, W = int(B), int(D), int(H), int(W)
    _need_cuda_f32(feats, "bev_pool feats")
    if coords.dtype != torch.int64 or not coords.is_cuda:
        raise RuntimeError("bev_pool: coords must be a CUDA int64 tensor [N,4]")
    feats = feats.contiguous()
    coords = coords.contiguous()
    N, Cc = feats.shape[0], feats.shape[1] if feats.dim() == 2 else 0
    if feats.dim() != 2 or coords.shape != (N, 4):
        raise RuntimeError(f"bev_pool: expected feats [N,C] and coords [N,4], got {tuple(feats.shape)}, {tuple(coords.shape)}")
    out = torch.empty((B, Cc, D, H, W), dtype=torch.float32, device=feats.device)
    wsb = int(lib.ss_bev_pool_workspace_bytes(max(N, 1), B * D * H * W))
    ws = torch.empty(wsb, dtype=torch.uint8, device=feats.device)
    rc = lib.ss_bev_pool_fwd(feats.data_ptr(), coords.data_ptr(), N, Cc, B, D, H, W, out.data_ptr(), ws.data_ptr(), wsb,
                             _stream())
    cabi.check(rc, "ss_bev_pool_fwd")
    return out


def ssc_confusion(pred: torch.Tensor, target: torch.Tensor, n_classes: int, nonempty: Optional[torch.Tensor] = None,
                  nonsurface: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
                  ignore_label: int = 255) -> torch.Tensor:
    """int64[C*C+3+C]: confusion matrix [target][prediction] of the remapped labels + completion (tp, fp, fn) + per-target count of
    predictions outside the class range;
    ``counts`` (zeroed by the caller) lets several samples accumulate into one tensor."""
    lib = cabi.load()
    if not pred.is_cuda or not target.is_cuda:
        raise RuntimeError("ssc_confusion: expected CUDA tensors (the hot path has no CPU fallback)")
    if pred.shape != target.shape:
        raise RuntimeError(f"ssc_confusion: shapes differ {tuple(pred.shape)} vs {tuple(target.shape)}")
    if pred.dtype != torch.uint8:
        pred = pred.to(torch.uint8)
    if target.dtype not in (torch.uint8, torch.int64):
        target = target.to(torch.int64)
    pred, target = pred.contiguous(), target.contiguous()

    def mask(m, name):
        if m is None:
            return None
        if m.shape != target.shape:
            raise RuntimeError(f"ssc_confusion: {name} has shape {tuple(m.shape)}, expected {tuple(target.shape)}")
        m = m.contiguous()
        return m.view(torch.uint8) if m.dtype == torch.bool else (m != 0).view(torch.uint8)

    ne, ns = mask(nonempty, "nonempty"), mask(nonsurface, "nonsurface")
    if counts is None:
        counts = torch.zeros(n_classes * n_classes + 3 + n_classes, dtype=torch.int64, device=pred.device)
    rc = lib.ss_ssc_confusion_fwd(pred.data_ptr(), target.data_ptr(), target.element_size(), _ptr(ne), _ptr(ns), pred.numel(),
                                  n_classes, ignore_label, counts.data_ptr(), _stream())
    cabi.check(rc, "ss_ssc_confusion_fwd")
    return counts


def deform_sample(x: torch.Tensor, offsets: torch.Tensor, groups: int, k: int, stride: int, pad: int, dil: int):
    """x [B,H,W,C] channels-last, offsets [B,2*k*k,Ho,Wo] NCHW -> S [B,Ho,Wo,groups,k*k,C/groups]."""
    lib = cabi.load()
    _need_cuda_f32(x, "deform_sample x"); _need_cuda_f32(offsets, "deform_sample offsets")
    if not (x.is_contiguous() and offsets.is_contiguous()):
        raise RuntimeError("deform_sample: inputs must be contiguous")
    B, H, W, Cc = x.shape
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    if tuple(offsets.shape) != (B, 2 * k * k, Ho, Wo):
        raise RuntimeError(f"deform_sample: offsets shape {tuple(offsets.shape)}, expected {(B, 2 * k * k, Ho, Wo)}")
    out = torch.empty((B, Ho, Wo, groups, k * k, Cc // groups), dtype=torch.float32, device=x.device)
    rc = lib.ss_deform_sample_fwd(x.data_ptr(), offsets.data_ptr(), out.data_ptr(), B, H, W, Cc, groups, k, k, stride, pad,
                                  dil, _stream())
    cabi.check(rc, "ss_deform_sample_fwd")
    return out


# ------------------------------------------------------------------------------------------
# resize + layout
# ------------------------------------------------------------------------------------------
def trilinear(x: torch.Tensor, size: Sequence[int], want_labels: bool = False):
    """x [B,Di,Hi,Wi,C] channels-last -> ([B,Do,Ho,Wo,C], labels uint8 [B,Do,Ho,Wo] or None)."""
    lib = cabi.load()
    _need_cuda_f32(x, "trilinear")
    if not x.is_contiguous():
        raise RuntimeError("trilinear: input must be contiguous channels-last")
    B, Di, Hi, Wi, Cc = x.shape
    Do, Ho, Wo = (int(s) for s in size)
    y = torch.empty((B, Do, Ho, Wo, Cc), dtype=torch.float32, device=x.device)
    labels = torch.empty((B, Do, Ho, Wo), dtype=torch.uint8, device=x.device) if want_labels else None
    rc = lib.ss_trilinear_fwd(x.data_ptr(), y.data_ptr(), _ptr(labels), B, Cc, Di, Hi, Wi, Do, Ho, Wo, _stream())
    cabi.check(rc, "ss_trilinear_fwd")
    return y, labels


def to_channels_last(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,C,*spatial] contiguous (NCHW / NCDHW) -> [B,*spatial,C] contiguous (``out``: contiguous destination)."""
    lib = cabi.load()
    _need_cuda_f32(x, "to_channels_last")
    x = x.contiguous()
    B, Cc = x.shape[0], x.shape[1]
    V = int(math.prod(x.shape[2:]))
    y = out if out is not None else torch.empty((B,) + tuple(x.shape[2:]) + (Cc,), dtype=torch.float32, device=x.device)
    if out is not None and (out.numel() != x.numel() or not out.is_contiguous()):
        raise RuntimeError("to_channels_last: out must be a contiguous tensor of the same size")
    cabi.check(lib.ss_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), B, Cc, V, Cc, _stream()), "ss_nchw_to_nhwc")
    return y


def to_channels_first(x: torch.Tensor) -> torch.Tensor:
    """[B,*spatial,C] contiguous -> [B,C,*spatial] contiguous."""
    lib = cabi.load()
    _need_cuda_f32(x, "to_channels_first")
    x = x.contiguous()
    B, Cc = x.shape[0], x.shape[-1]
    V = int(math.prod(x.shape[1:-1]))
    y = torch.empty((B, Cc) + tuple(x.shape[1:-1]), dtype=torch.float32, device=x.device)
    cabi.check(lib.ss_nhwc_to_nchw(x.data_ptr(), y.data_ptr(), B, Cc, V, Cc, _stream()), "ss_nhwc_to_nchw")
    return y
