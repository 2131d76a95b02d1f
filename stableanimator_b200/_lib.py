"""ctypes binding of ``libsa_b200.so`` (the C ABI declared in ``include/sa_b200.h``).

The product path has no CPU / PyTorch fallback: if the library is missing or a call fails the
binding raises (``SaLibraryError`` / ``SaKernelError``) instead of computing anything elsewhere.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libsa_b200.so"


class SaLibraryError(RuntimeError):
    pass


class SaKernelError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of ``sa_gemm_args`` (include/sa_b200.h)."""

    _fields_ = [
        ("a0", C.c_void_p), ("a1", C.c_void_p),
        ("c0", C.c_int32), ("c1", C.c_int32),
        ("lda0", C.c_int32), ("lda1", C.c_int32),
        ("d1", C.c_int32), ("d2", C.c_int32), ("d3", C.c_int32),
        ("ntaps", C.c_int32),
        ("tap_off", C.c_int32 * 27),
        ("w", C.c_void_p),
        ("n", C.c_int32),
        ("bias", C.c_void_p), ("row_bias", C.c_void_p), ("ldrb", C.c_int32), ("row_scale", C.c_void_p),
        ("rows_per_group", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int32),
        ("blend_x", C.c_void_p), ("ldx", C.c_int32),
        ("alpha", C.c_float),
        ("geglu", C.c_int32),
        ("out", C.c_void_p), ("ldc", C.c_int32),
        ("out_f32", C.c_int32),
        ("tile_n", C.c_int32),
        ("out_scale", C.c_float),
        ("f16_from_col", C.c_int32),
        ("gn_stats", C.c_void_p), ("gn_cpg", C.c_int32), ("gn_rows_per_stat", C.c_int32),
        ("ldc2", C.c_int64), ("ldc3", C.c_int64),
        ("in_stride", C.c_int32),
        ("act_f16", C.c_int32),
        ("gn_stats_zeroed", C.c_int32),
    ]


class FfArgs(C.Structure):
    """Mirror of ``sa_ff_args`` (include/sa_b200.h)."""

    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int32), ("m", C.c_int64), ("c", C.c_int32),
        ("addv", C.c_void_p), ("rows_per_group", C.c_int32),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("blend_x", C.c_void_p), ("ldbx", C.c_int32), ("alpha", C.c_float),
        ("out", C.c_void_p), ("ldc", C.c_int32), ("act_f16", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the CUDA library; raises SaLibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SaLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m stableanimator_b200.build` "
            "(there is no CPU fallback for the hot path)")
    try:
        l = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise SaLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    l.sa_version.restype = C.c_int
    l.sa_last_error.restype = C.c_char_p
    _declare(l)
    _lib = l
    return l


# name -> argtypes; every function returns int (0 = ok)
_P, _I, _F, _L = C.c_void_p, C.c_int32, C.c_float, C.c_int64
SIGNATURES: dict[str, list] = {
    "sa_gemm_tap": [C.POINTER(GemmArgs), _P],
    "sa_gemm_pick_tile": [_I, _I, _I],
    "sa_ff_fused": [C.POINTER(FfArgs), _P],
    "sa_attn_spatial": [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "sa_attn_temporal": [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "sa_attn_id": [_P, _I, _P, _P, _L, _L, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P],
    "sa_attn_id_finalize": [_P, _P, _L, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "sa_groupnorm": [_P, _P, _I, _I, _I, _I, _L, _I, _P, _P, _F, _I, _P, _P, _I, _I, _P],
    "sa_groupnorm_apply": [_P, _P, _I, _I, _I, _I, _L, _I, _P, _P, _F, _I, _P, _P, _I, _I, _P],
    "sa_layernorm": [_P, _I, _L, _I, _P, _I, _P, _I, _P, _P, _F, _P, _I, _I, _P],
    "sa_small_linear": [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _P, _I, _P],
    "sa_sinusoid": [_P, _I, _I, _P, _P],
    "sa_softmax_rows": [_P, _L, _I, _I, _F, _P, _L, _P],
    "sa_time_conv_out": [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "sa_layernorm_f32": [_P, _I, _I, _I, _P, _P, _F, _P, _I, _P],
    "sa_perceiver_attn": [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P],
    "sa_conv_direct": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _I, _P, _I, _P],
    "sa_upsample2x": [_P, _I, _I, _I, _I, _P, _P],
    "sa_im2col_s2": [_P, _I, _I, _I, _I, _P, _P],
    "sa_im2col_s2_pad": [_P, _I, _I, _I, _I, _I, _P, _P],
    "sa_im2col3x3_c8": [_P, _I, _I, _I, _P, _P],
    "sa_nchw_to_nhwc": [_P, _I, _I, _I, _I, _P, _I, _P],
    "sa_nhwc_to_nchw_f32": [_P, _I, _I, _I, _I, _P, _P],
    "sa_sampler_prepare": [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _I, _P],
    "sa_sampler_accumulate": [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "sa_sampler_step": [_P, _P, _P, _I, _I, _F, _F, _P, _P],
    "sa_gauss_blur_axis": [_P, _P, _I, _I, _I, _P, _I, _I, _F, _F, _P],
    "sa_resize_bicubic_norm": [_P, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P],
    "sa_patchify": [_P, _I, _I, _I, _I, _P, _I, _P],
    "sa_attn_tokens": [_P, _I, _I, _I, _I, _I, _F, _P, _I, _P],
}


def _declare(l: C.CDLL) -> None:
    for name, argtypes in SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int


# kernels launched by one successful call of each entry point (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {"sa_groupnorm": 2, "sa_groupnorm_apply": 1}
launch_count = 0


def check(rc: int, what: str) -> None:
    global launch_count
    launch_count += KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        msg = lib().sa_last_error().decode("utf-8", "replace")
        raise SaKernelError(f"{what} failed with code {rc}: {msg}")
