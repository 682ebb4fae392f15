"""ctypes binding of libdiffute_b200.so (the C-ABI declared in include/diffute_b200.h).

The library is the only compute path: if it is missing, or a call fails, we raise — there is no
CPU / PyTorch fallback anywhere in diffute_b200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffute_b200.so")


class DfuError(RuntimeError):
    pass


class GemmOperand(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_mode", C.c_int32), ("a_rows", C.c_int32), ("a_ld", C.c_int32),
        ("a_h", C.c_int32), ("a_w", C.c_int32), ("a_c", C.c_int32), ("a_plane", C.c_int32),
        ("b", C.c_void_p), ("b_rows", C.c_int32), ("b_ld", C.c_int32), ("b_plane", C.c_int32),
        ("ntaps", C.c_int32), ("k_per_tap", C.c_int32),
        ("tap_dn", C.c_int8 * 9), ("tap_dy", C.c_int8 * 9), ("tap_dx", C.c_int8 * 9), ("b_static", C.c_int8),
        ("_pad", C.c_int8 * 4),
    ]


class Gemm(C.Structure):
    _fields_ = [
        ("m", C.c_int32), ("n", C.c_int32), ("ngroups", C.c_int32), ("npass", C.c_int32),
        ("g", GemmOperand * 2),
        ("batch", C.c_int32), ("a_batch_rows", C.c_int32), ("b_batch_rows", C.c_int32),
        ("conv", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("epi", C.c_int32), ("act", C.c_int32), ("alpha", C.c_float),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rowvec_ld", C.c_int32), ("rows_per_sample", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int32),
        ("out_f32", C.c_void_p), ("ldo", C.c_int32),
        ("out_f16", C.c_void_p), ("ldh", C.c_int32), ("out_planes", C.c_int32), ("out_plane_stride", C.c_int64),
        ("block_n", C.c_int32), ("splits", C.c_int32), ("stages", C.c_int32), ("kernel", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("sync_words", C.c_void_p),
    ]


class PackJob(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src2", C.c_void_p), ("dst", C.c_void_p), ("plane_stride", C.c_int64),
        ("rows", C.c_int32), ("cin", C.c_int32), ("taps", C.c_int32), ("mode", C.c_int32),
        ("dst_row0", C.c_int32), ("dst_ld", C.c_int32), ("geglu", C.c_int32), ("planes", C.c_int32),
    ]


EPI_F32, EPI_F16, EPI_GEGLU = 0, 1, 2

_lib = None


def lib() -> C.CDLL:
    """Load (building first if the sources changed and nvcc is available) the native library."""
    global _lib
    if _lib is not None:
        return _lib
    from . import _build
    # DFU_TRACE=1 loads the diagnostic build with in-kernel timeline records (diffute_b200/trace.py); never the default
    path = _build.build(trace=os.environ.get("DFU_TRACE") == "1")
    if not os.path.exists(path):
        raise DfuError(f"{path} missing: the CUDA extension is required (no fallback)")
    L = C.CDLL(path)
    L.dfu_version.restype = C.c_int
    L.dfu_last_error.restype = C.c_char_p
    L.dfu_num_sms.restype = C.c_int
    L.dfu_gemm.argtypes = [C.POINTER(Gemm), C.c_void_p]
    L.dfu_gemm.restype = C.c_int
    L.dfu_gemm_stats.argtypes = [C.POINTER(C.c_int64)]
    L.dfu_gemm_stats.restype = None
    L.dfu_gemm_plan.argtypes = [C.POINTER(Gemm), C.POINTER(C.c_int32)]
    L.dfu_gemm_plan.restype = C.c_int
    L.dfu_gemm_workspace.argtypes = [C.POINTER(Gemm)]
    L.dfu_gemm_workspace.restype = C.c_size_t
    _declare_rest(L)
    _lib = L
    return L


def _declare_rest(L):
    """argtypes for the non-GEMM entry points (kept in one table so tests can check header <-> binding)."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args


# name -> (restype, argtypes); filled as entry points are added (see include/diffute_b200.h)
_vp, _i, _f, _i64, _sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t
SIGNATURES = {
    "dfu_groupnorm_workspace": (_sz, [_i, _i, _i, _i]),
    "dfu_groupnorm": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _f, _i, _vp, _i, _i64, _vp, _vp, _vp, _sz, _vp,
                           _vp]),
    "dfu_layernorm": (_i, [_vp, _i, _i, _vp, _vp, _f, _vp, _i, _i64, _vp, _vp]),
    "dfu_patchify_f16": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i64, _vp]),
    "dfu_cast_f16": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i64, _vp]),
    "dfu_timestep_embedding": (_i, [_vp, _i, _i, _i, _f, _vp, _vp]),
    "dfu_gemv": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "dfu_conv_small_in": (_i, [_vp, _i, _i64, _vp, _i, _i64, _vp, _i, _i64, _i, _i, _i, _i, _i, _vp, _vp, _i, _f, _vp,
                               _vp]),
    "dfu_conv_small_out": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dfu_philox_normal": (_i, [C.c_uint64, C.c_uint32, _i64, _vp, _vp, _vp]),
    "dfu_glue_preprocess": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "dfu_glyph_preprocess_workspace": (_sz, [_i, _i, _i]),
    "dfu_glyph_preprocess": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "dfu_glue_composite": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "dfu_axpbypcz": (_i, [_vp, _vp, _vp, _f, _f, _f, _vp, _i64, _vp]),
    "dfu_scheduler_step": (_i, [_vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _i, _vp, _vp, _i64, _vp]),
    "dfu_axpby_rows": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "dfu_gaussian_sample": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "dfu_softmax_rows": (_i, [_vp, _i, _i, _i, _f, _vp, _i, _i, _i64, _vp]),
    "dfu_attention_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "dfu_attention_plan": (_i, [_i, _i, _i, _i, _i, C.POINTER(C.c_int32)]),
    "dfu_attention": (_i, [_vp, _i, _i, _i64, _vp, _i, _i, _vp, _i, _i, _i64, _i, _i, _i, _i, _i, _f, _vp, _i, _i64,
                           _i, _vp, _sz, _vp]),
    "dfu_transpose_f16": (_i, [_vp, _i, _i, _i, _i, _i64, _vp, _i64, _vp]),
    "dfu_pack_weights": (_i, [_vp, _vp, _i, _i64, _vp]),
    "dfu_trace_set_gemm": (_i, [_vp]),
    "dfu_trace_set_gemm2": (_i, [_vp]),
    "dfu_trace_set_attn": (_i, [_vp]),
    "dfu_trace_set_norm": (_i, [_vp]),
    "dfu_trace_set_misc": (_i, [_vp]),
}


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dfu_last_error().decode(errors="replace")
        raise DfuError(f"{what} failed (rc={rc}): {msg}")
