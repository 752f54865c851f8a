"""ctypes binding of include/stereoscene_b200.h.

The product path has no fallback: if ``libstereoscene_b200.so`` is missing or its ABI version
does not match, importing an op raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

ABI_VERSION = 6

SS_ACT_NONE, SS_ACT_RELU, SS_ACT_GELU, SS_ACT_SWISH, SS_ACT_SIGMOID = 0, 1, 2, 3, 4
SS_MATH_TF32, SS_MATH_3XTF32, SS_MATH_TF32X3, SS_MATH_F16X3, SS_MATH_F16 = 0, 1, 2, 3, 4


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "B", "Din", "Hin", "Win", "Cin", "Dout", "Hout", "Wout", "Cout",
        "kd", "kh", "kw", "sd", "sh", "sw", "pd", "ph", "pw", "dd", "dh", "dw",
        "transposed", "in_ldc", "out_ldc", "in_act", "out_act", "math", "cout_packed", "stats_d0", "stats_d1")] + [("acc_scale", C.c_float), ("accumulate", C.c_int32),
                                                                                                      ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64)]


class ConvJoin(C.Structure):
    _fields_ = [("out_scale", C.c_void_p), ("out_shift", C.c_void_p), ("res", C.c_void_p), ("res_scale", C.c_void_p),
                ("res_shift", C.c_void_p), ("res_ldc", C.c_int32), ("res_act", C.c_int32)]


class NativeLibraryError(RuntimeError):
    pass


_vp, _i, _ll, _d, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_float, C.c_size_t

# name -> (restype, argtypes); exactly the declarations of include/stereoscene_b200.h
SIGNATURES = {
    "ss_abi_version": (_i, []),
    "ss_last_error_string": (C.c_char_p, []),
    "ss_launch_count": (_ll, []),
    "ss_kernel_census": (_i, [C.c_char_p, _sz]),
    "ss_conv3d_fwd": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ss_conv3d_tc_fwd": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ss_conv3d_tc_join_supported": (_i, [C.POINTER(ConvDesc)]),
    "ss_conv3d_tc_f16x3_supported": (_i, [C.POINTER(ConvDesc)]),
    "ss_conv3d_tc_join_fwd": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, C.POINTER(ConvJoin), _vp, _vp]),
    "ss_gn_finalize": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _f, _vp, _vp, _i, _vp]),
    "ss_gn_finalize_gated": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ss_aspp_pool_shift": (_i, [_vp, _d, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ss_ca3d_gate": (_i, [_vp, _d, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ss_affine_join_fwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _ll, _i, _i, _i, _i, _vp, _vp]),
    "ss_channel_sums_fwd": (_i, [_vp, _vp, _vp, _i, _i, _ll, _i, _i, _vp, _vp]),
    "ss_stem_conv2d_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ss_dwconv2d_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ss_se_fc_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, C.c_float, _i, _vp]),
    "ss_softmax_d_fwd": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "ss_gwc_warp_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ss_bri_workspace_bytes": (_sz, [_i, _i, _i]),
    "ss_bri_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _i, _i, _i, _vp]),
    "ss_splat_index_workspace_bytes": (_sz, [_ll]),
    "ss_splat_build_index": (_i, [_vp, C.POINTER(_f), C.POINTER(_f), _i, _i, _i, _i, _ll, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ss_lift_splat_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ss_bev_pool_workspace_bytes": (_sz, [_ll, _ll]),
    "ss_bev_pool_fwd": (_i, [_vp, _vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ss_trilinear_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ss_ssc_confusion_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _ll, _i, _i, _vp, _vp]),
    "ss_deform_sample_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ss_peer_pool_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "ss_peer_pool_free": (_i, [_vp]),
    "ss_peer_ipc_export": (_i, [_vp, _vp]),
    "ss_peer_ipc_open": (_i, [_vp, _i, C.POINTER(_vp)]),
    "ss_peer_ipc_close": (_i, [_vp]),
    "ss_peer_epoch_bump": (_i, [_vp, _vp]),
    "ss_peer_halo_push": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ss_peer_stats_allreduce": (_i, [_vp, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp]),
    "ss_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _ll, _i, _vp]),
    "ss_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _ll, _i, _vp]),
}

_lib = None


def load(path: str | None = None):
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("STEREOSCENE_B200_LIB") or _build.library_path()
    if not os.path.exists(path):
        raise NativeLibraryError(
            f"{path} not found: the CUDA extension is not built. Run `python -m stereoscene_b200._build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{path} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.ss_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI mismatch: library {lib.ss_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().ss_last_error_string().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().ss_launch_count())


def kernel_census() -> dict:
    """kernel name -> launches issued by this process so far (per kernel family of the library)."""
    buf = C.create_string_buffer(8192)
    load().ss_kernel_census(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        k, _, v = line.partition("=")
        out[k] = int(v)
    return out
