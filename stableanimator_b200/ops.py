"""Python-side launchers: torch tensors in, raw pointers through the C ABI, torch tensors out.

PyTorch is used for device memory and streams only -- every arithmetic op below runs in
``libsa_b200.so``.  Layout convention: activations are bf16 and channel-innermost
(``[frames, h, w, C]`` == ``[frames*h*w, C]``).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import functools
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import FfArgs, GemmArgs, check, lib

TAPS_1 = ((0, 0, 0),)
# 3x3 conv, pad 1: tap index = ky*3 + kx, offset (dx, dy, 0) = (kx-1, ky-1, 0)
TAPS_3X3 = tuple((kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3))
# temporal conv k=(3,1,1), pad (1,0,0) on a (S, F, B) grid: offset along dim 2 (frames)
TAPS_T3 = ((0, -1, 0), (0, 0, 0), (0, 1, 0))


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_of(t: torch.Tensor):
    """Context that makes ``t``'s device current for a launch (kernels run on the CURRENT device's stream, and
    ``cudaFuncSetAttribute`` / the SM count are per device): free when it already is, so a model built on ``cuda:1``
    works while ``cuda:0`` is current instead of faulting on foreign pointers."""
    if not t.is_cuda:
        raise ValueError(f"expected a CUDA tensor, got one on {t.device} (the hot path has no CPU fallback)")
    if t.device.index == torch.cuda.current_device():
        return contextlib.nullcontext()
    return torch.cuda.device(t.device)


# NVTX range per kernel family (tracing): SA_NVTX=1, e.g. under `nsys` / `ncu --nvtx`
_NVTX = os.environ.get("SA_NVTX", "0") == "1"


def _launcher(family: str):
    """Wrap a launcher: first tensor argument's device becomes current; optional NVTX range named after the family."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(t0, *args, **kwargs):
            with device_of(t0):
                if not _NVTX:
                    return fn(t0, *args, **kwargs)
                torch.cuda.nvtx.range_push(family)
                try:
                    return fn(t0, *args, **kwargs)
                finally:
                    torch.cuda.nvtx.range_pop()
        wrapped.family = family
        return wrapped
    return deco


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


ACT_DTYPES = (torch.bfloat16, torch.float16)


def _chk_bf16(t: torch.Tensor, name: str, like: Optional[torch.Tensor] = None) -> None:
    """A 16-bit activation / weight tensor: CUDA, bf16 (default mode) or fp16 (the reference's default inference dtype,
    inference_basic.py:486), channel-innermost; ``like``: must share that tensor's format."""
    if t.dtype not in ACT_DTYPES or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA bf16 / fp16 tensor, got {t.dtype} on {t.device}")
    if like is not None and t.dtype != like.dtype:
        raise ValueError(f"{name} is {t.dtype} but the call runs in {like.dtype}: all 16-bit operands share one format")
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must be channel-innermost (stride(-1) == 1)")


def _f16(t: torch.Tensor) -> int:
    return int(t.dtype == torch.float16)


def _chk_f32(t: Optional[torch.Tensor], name: str) -> None:
    if t is None:
        return
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA fp32 tensor")


def pick_tile(m: int, n: int, geglu: bool = False) -> int:
    return lib().sa_gemm_pick_tile(m, n, int(geglu))


@_launcher("gemm_tap")
def gemm_tap(
    a0: torch.Tensor,
    w: torch.Tensor,
    *,
    a1: Optional[torch.Tensor] = None,
    grid: Optional[Sequence[int]] = None,
    taps: Sequence[Sequence[int]] = TAPS_1,
    n: Optional[int] = None,
    bias: Optional[torch.Tensor] = None,
    row_bias: Optional[torch.Tensor] = None,
    row_scale: Optional[torch.Tensor] = None,
    rows_per_group: int = 1,
    residual: Optional[torch.Tensor] = None,
    blend_x: Optional[torch.Tensor] = None,
    alpha: float = 0.0,
    geglu: bool = False,
    out: Optional[torch.Tensor] = None,
    out_f32: bool = False,
    tile_n: int = 0,
    out_scale: float = 1.0,
    f16_from_col: int = 0,
    gn_stats: Optional[torch.Tensor] = None,
    gn_rows_per_stat: int = 0,
    in_stride: int = 1,
    gn_stats_zeroed: bool = False,
) -> torch.Tensor:
    """``out = epilogue(sum_taps A(shifted) @ W^T)`` -- see ``sa_gemm_tap`` in include/sa_b200.h.

    ``a0`` / ``a1``: bf16, ``[..., C]`` with a uniform row stride (``stride(-2)``), all leading dims
    dense over the position grid ``grid = (d1, d2, d3)`` (default ``(rows, 1, 1)``).
    ``w``: bf16 ``[N, ntaps*(C0+C1)]``.
    """
    _chk_bf16(a0, "a0")
    _chk_bf16(w, "w", a0)
    c0 = a0.shape[-1]
    lda0 = a0.stride(-2) if a0.dim() >= 2 else c0
    src_rows = a0.numel() // c0
    if in_stride not in (1, 2) or (in_stride == 2 and (grid is None or src_rows != 4 * grid[0] * grid[1] * grid[2])):
        raise ValueError("in_stride=2 needs an explicit grid (d1, d2, d3) and sources of 2*d1 x 2*d2 x d3 positions")
    rows = src_rows // (in_stride * in_stride)   # output positions
    c1, lda1 = 0, 0
    if a1 is not None:
        _chk_bf16(a1, "a1", a0)
        c1 = a1.shape[-1]
        lda1 = a1.stride(-2)
        if a1.numel() // c1 != src_rows:
            raise ValueError("a0 and a1 must cover the same positions")
    if grid is None:
        grid = (rows, 1, 1)
    d1, d2, d3 = (int(x) for x in grid)
    if d1 * d2 * d3 != rows:
        raise ValueError(f"grid {grid} does not match {rows} positions")
    ntaps = len(taps)
    n_mma = int(n) if n is not None else w.shape[0]
    if w.dim() != 2 or w.shape[1] != ntaps * (c0 + c1):
        raise ValueError(f"w must be [N, K] with K={ntaps}*({c0}+{c1}), got {tuple(w.shape)}")
    if w.stride(0) != w.shape[1]:   # the W tensor map is built with a row pitch of K elements
        raise ValueError("w must have dense rows (stride(0) == K); pass .contiguous() for a sliced / padded weight")
    if geglu and tile_n and n_mma % tile_n:
        raise ValueError("tile_n must divide n for GEGLU-interleaved weights")
    n_out = n_mma // 2 if geglu else n_mma
    if out is None:
        out = torch.empty((rows, n_out), device=a0.device,
                          dtype=torch.float32 if out_f32 else a0.dtype)
    elif out.dtype != torch.float32 and out.dtype != a0.dtype:
        raise ValueError(f"out is {out.dtype} but the call runs in {a0.dtype}")
    _chk_f32(bias, "bias")
    _chk_f32(row_scale, "row_scale")
    args = GemmArgs()
    args.a0, args.a1 = a0.data_ptr(), _ptr(a1)
    args.c0, args.c1, args.lda0, args.lda1 = c0, c1, lda0, lda1
    args.d1, args.d2, args.d3 = d1, d2, d3
    args.ntaps = ntaps
    for i, t in enumerate(taps):
        args.tap_off[3 * i], args.tap_off[3 * i + 1], args.tap_off[3 * i + 2] = t
    args.w, args.n = w.data_ptr(), n_mma
    args.bias, args.row_scale = _ptr(bias), _ptr(row_scale)
    if row_bias is not None:
        if row_bias.dtype != torch.float32 or row_bias.stride(-1) != 1:
            raise ValueError("row_bias must be fp32 with unit inner stride")
        args.row_bias, args.ldrb = row_bias.data_ptr(), row_bias.stride(0)
    args.rows_per_group = rows_per_group
    if residual is not None:
        _chk_bf16(residual, "residual", a0)
        args.residual, args.ldr = residual.data_ptr(), residual.stride(-2)
    if blend_x is not None:
        _chk_bf16(blend_x, "blend_x", a0)
        args.blend_x, args.ldx = blend_x.data_ptr(), blend_x.stride(-2)
    args.alpha = float(alpha)
    args.act_f16 = _f16(a0)
    args.geglu = int(geglu)
    args.out, args.ldc = out.data_ptr(), out.stride(-2)
    if out.dim() == 4 and not out.is_contiguous():
        # strided output grid (every other pixel of a larger image): [d3, d2, d1, n_out] view with arbitrary strides
        if tuple(out.shape[:3]) != (d3, d2, d1) or out.stride(-1) != 1:
            raise ValueError("a strided `out` must be a [d3, d2, d1, n_out] view")
        args.ldc2, args.ldc3 = out.stride(1), out.stride(0)
    args.out_f32 = int(out.dtype == torch.float32)
    args.tile_n = tile_n
    args.out_scale = float(out_scale)
    args.f16_from_col = int(f16_from_col)
    args.in_stride = int(in_stride)
    if gn_stats is not None:
        # fused statistics of the output for the GroupNorm(32) that consumes it: fp64 [rows / rows_per_stat, 32, 2]
        if gn_stats.dtype != torch.float64 or not gn_stats.is_contiguous() or n_out % 32 or \
                gn_stats.numel() != rows // gn_rows_per_stat * 64:
            raise ValueError("gn_stats must be a contiguous fp64 [rows / gn_rows_per_stat, 32, 2] tensor")
        args.gn_stats, args.gn_cpg, args.gn_rows_per_stat = gn_stats.data_ptr(), n_out // 32, int(gn_rows_per_stat)
        args.gn_stats_zeroed = int(bool(gn_stats_zeroed))  # the caller filled it with zeros: no memset in front of the kernel
    check(lib().sa_gemm_tap(C.byref(args), _stream()), "sa_gemm_tap")
    return out


FF_FUSED_CHANNELS = (64, 128, 256, 320)


@_launcher("ff_fused")
def ff_fused(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor,
             w2: torch.Tensor, b2: torch.Tensor, *, addv: Optional[torch.Tensor] = None, rows_per_group: int = 1,
             blend_x: Optional[torch.Tensor] = None, alpha: float = 0.0, eps: float = 1e-5,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out = W2 . GEGLU(W1 . LN(x + addv) + b1) + b2 + (x + addv)`` (optionally alpha-blended with ``blend_x``) in ONE
    kernel -- see ``sa_ff_fused`` in include/sa_b200.h.  ``w1`` / ``b1`` are GEGLU-interleaved (:func:`interleave_geglu`)."""
    _chk_bf16(x, "x")
    _chk_bf16(w1, "w1", x)
    _chk_bf16(w2, "w2", x)
    c = x.shape[-1]
    rows = x.numel() // c
    if c not in FF_FUSED_CHANNELS:
        raise ValueError(f"ff_fused supports C in {FF_FUSED_CHANNELS}, got {c}")
    if tuple(w1.shape) != (8 * c, c) or tuple(w2.shape) != (c, 4 * c) or not w1.is_contiguous() or not w2.is_contiguous():
        raise ValueError("w1 must be contiguous [8C, C] (GEGLU-interleaved) and w2 contiguous [C, 4C]")
    for t, nm in ((gamma, "gamma"), (beta, "beta"), (b1, "b1"), (b2, "b2"), (addv, "addv")):
        _chk_f32(t, nm)
    if out is None:
        out = torch.empty((rows, c), device=x.device, dtype=x.dtype)
    a = FfArgs()
    a.x, a.ldx, a.m, a.c = x.data_ptr(), x.stride(-2), rows, c
    a.addv, a.rows_per_group = _ptr(addv), int(rows_per_group)
    a.gamma, a.beta, a.eps = gamma.data_ptr(), beta.data_ptr(), float(eps)
    a.w1, a.b1, a.w2, a.b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
    if blend_x is not None:
        _chk_bf16(blend_x, "blend_x", x)
        a.blend_x, a.ldbx = blend_x.data_ptr(), blend_x.stride(-2)
    a.alpha = float(alpha)
    a.act_f16 = _f16(x)
    a.out, a.ldc = out.data_ptr(), out.stride(-2)
    check(lib().sa_ff_fused(C.byref(a), _stream()), "sa_ff_fused")
    return out


# (py, px, taps) of the four output phases of nearest-x2 + conv3x3, in the order upsample_conv_phases() packs them
UPSAMPLE_PHASE_TAPS = tuple((py, px, tuple((px - 1 + tx, py - 1 + ty, 0) for ty in (0, 1) for tx in (0, 1)))
                            for py in (0, 1) for px in (0, 1))


def upsample_conv_phases(w_oihw: torch.Tensor):
    """Upsample2D = nearest x2 then Conv2d 3x3 pad 1 (unet.py:1789-1807) as four 2x2-tap convolutions on the
    LOW-resolution input, one per output phase (py, px): out[2y + py, 2x + px] = sum over (ty, tx) in {0,1}^2 of
    Wp[ty, tx] . in[y + py - 1 + ty, x + px - 1 + tx], where Wp sums the original taps that read the same input pixel
    (py = 0: rows {ky0} | {ky1 + ky2}; py = 1: rows {ky0 + ky1} | {ky2}; same along x).  4/9 of the multiply-adds and
    no upsampled tensor.  Returns [(py, px, taps, weight fp32 [O, (ty, tx, ci)])]; the zero padding of the upsampled
    image maps onto the zero padding of the low-resolution one."""
    w = w_oihw.float()
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    out = []
    for py in (0, 1):
        for px in (0, 1):
            blocks, taps = [], []
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = sum(w[:, :, ky, kx] for ky in rows[py][ty] for kx in rows[px][tx])   # [O, I]
                    blocks.append(acc)
                    taps.append((px - 1 + tx, py - 1 + ty, 0))
            out.append((py, px, tuple(taps), torch.cat(blocks, dim=1).contiguous()))
    return out


def interleave_geglu(w: torch.Tensor, b: Optional[torch.Tensor]):
    """Reorder GEGLU projection rows ``[value(4C) ; gate(4C)]`` (unet.py:865-870) into 64-wide
    blocks ``[v0..63 | g0..63 | v64..127 | g64..127 | ...]`` so one MMA tile holds both halves."""
    n2 = w.shape[0]
    h = n2 // 2
    assert h % 64 == 0
    idx = torch.arange(n2, device=w.device).reshape(2, h // 64, 64).permute(1, 0, 2).reshape(-1)
    return w[idx].contiguous(), (None if b is None else b[idx].contiguous())


# ----------------------------------------------------------------------------- attention
@_launcher("attn_spatial")
def attn_spatial(qkv: torch.Tensor, frames: int, s: int, heads: int,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T / 8) v per (frame, head); qkv 16-bit [frames*s, 3C] -> bf16 [frames*s, C].
    q | k are bf16, the v block holds IEEE fp16 bit patterns (``gemm_tap(..., f16_from_col=2C)`` or
    :func:`v_block_to_f16_`)."""
    _chk_bf16(qkv, "qkv")
    c = heads * 64
    if out is None:
        out = torch.empty((frames * s, c), device=qkv.device, dtype=qkv.dtype)
    check(lib().sa_attn_spatial(qkv.data_ptr(), qkv.stride(-2), frames, s, c, heads, out.data_ptr(),
                                out.stride(-2), _f16(qkv), _stream()), "sa_attn_spatial")
    return out


def v_block_to_f16_(qkv: torch.Tensor, c: int) -> torch.Tensor:
    """In place: re-encode the v block (columns >= 2C) of an all-bf16 q|k|v tensor as fp16 bit patterns."""
    if qkv.dtype == torch.float16:
        return qkv
    v = qkv[:, 2 * c:3 * c]
    v.copy_(v.float().clamp(-65504.0, 65504.0).half().view(torch.bfloat16))
    return qkv


@_launcher("attn_temporal")
def attn_temporal(qkv: torch.Tensor, b: int, f: int, s: int, heads: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk_bf16(qkv, "qkv")
    c = heads * 64
    if out is None:
        out = torch.empty((b * f * s, c), device=qkv.device, dtype=qkv.dtype)
    check(lib().sa_attn_temporal(qkv.data_ptr(), qkv.stride(-2), b, f, s, c, heads, out.data_ptr(),
                                 out.stride(-2), _f16(qkv), _stream()), "sa_attn_temporal")
    return out


@_launcher("attn_id")
def attn_id(q: torch.Tensor, ik: torch.Tensor, iv: torch.Tensor, b: int, f: int, s: int, heads: int,
            stats: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ip = softmax(q ik^T / 8) iv over the 4 face tokens; stats (fp64 [b*f, 2]) gets Sum, Sum^2.
    ik / iv: fp32 views [b, 4, C] (any batch/token strides, unit channel stride)."""
    _chk_bf16(q, "q")
    c = heads * 64
    for t in (ik, iv):
        if t.dtype != torch.float32 or t.shape != (b, 4, c) or t.stride(2) != 1:
            raise ValueError("ik/iv must be fp32 [b, 4, C] with unit channel stride")
    if ik.stride() != iv.stride():
        raise ValueError("ik and iv must share strides")
    if out is None:
        out = torch.empty((b * f * s, c), device=q.device, dtype=q.dtype)
    check(lib().sa_attn_id(q.data_ptr(), q.stride(-2), ik.data_ptr(), iv.data_ptr(), ik.stride(1), ik.stride(0),
                           b, f, s, c, heads, 4, out.data_ptr(), out.stride(-2), stats.data_ptr(), _f16(q), _stream()),
          "sa_attn_id")
    return out


@_launcher("attn_id")
def attn_id_finalize(stats, vclip, wv, w_rowsum, bias, b, f, s, c, row_scale, row_bias) -> None:
    """vclip: fp32 view [b, C] (any batch stride); the rest contiguous fp32."""
    for t, nm in ((wv, "wv"), (w_rowsum, "w_rowsum"), (bias, "bias"),
                  (row_scale, "row_scale"), (row_bias, "row_bias")):
        _chk_f32(t, nm)
    if vclip.dtype != torch.float32 or vclip.stride(1) != 1:
        raise ValueError("vclip must be fp32 with unit channel stride")
    check(lib().sa_attn_id_finalize(stats.data_ptr(), vclip.data_ptr(), vclip.stride(0), wv.data_ptr(),
                                    w_rowsum.data_ptr(), bias.data_ptr(), b, f, s, c, row_scale.data_ptr(),
                                    row_bias.data_ptr(), _stream()), "sa_attn_id_finalize")


# ----------------------------------------------------------------------------- normalisation
@_launcher("groupnorm")
def groupnorm(x0: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, rows_per_stat: int,
              *, x1: Optional[torch.Tensor] = None, silu: bool = True, stats_ws: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, stats_ready: bool = False) -> torch.Tensor:
    """``stats_ready``: ``stats_ws`` was filled by the epilogue of the ``gemm_tap(..., gn_stats=stats_ws)`` launch that
    produced ``x0`` -- only the normalise (+SiLU) pass runs (``sa_groupnorm_apply``)."""
    _chk_bf16(x0, "x0")
    c0 = x0.shape[-1]
    rows = x0.numel() // c0
    c1, ld1 = 0, 0
    if x1 is not None:
        _chk_bf16(x1, "x1", x0)
        c1, ld1 = x1.shape[-1], x1.stride(-2)
    ns = rows // rows_per_stat
    if ns * rows_per_stat != rows:
        raise ValueError("rows must be a multiple of rows_per_stat")
    if stats_ready and (stats_ws is None or x1 is not None or stats_ws.numel() != ns * 64):
        raise ValueError("stats_ready needs the fp64 [ns, 32, 2] statistics of a single-source input")
    if stats_ws is None:
        stats_ws = torch.empty((ns, 32, 2), device=x0.device, dtype=torch.float64)
    if out is None:
        out = torch.empty((rows, c0 + c1), device=x0.device, dtype=x0.dtype)
    _chk_f32(gamma, "gamma")
    _chk_f32(beta, "beta")
    fn, what = (lib().sa_groupnorm_apply, "sa_groupnorm_apply") if stats_ready else (lib().sa_groupnorm, "sa_groupnorm")
    check(fn(x0.data_ptr(), _ptr(x1), c0, c1, x0.stride(-2), ld1, rows_per_stat, ns,
             gamma.data_ptr(), beta.data_ptr(), float(eps), int(silu), stats_ws.data_ptr(),
             out.data_ptr(), out.stride(-2), _f16(x0), _stream()), what)
    return out


@_launcher("layernorm")
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, addv: Optional[torch.Tensor] = None,
              rows_per_group: int = 1, x_out: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, eps: float = 1e-5) -> torch.Tensor:
    _chk_bf16(x, "x")
    c = x.shape[-1]
    rows = x.numel() // c
    if out is None:
        out = torch.empty((rows, c), device=x.device, dtype=x.dtype)
    _chk_f32(addv, "addv")
    check(lib().sa_layernorm(x.data_ptr(), x.stride(-2), rows, c, _ptr(addv), rows_per_group, _ptr(x_out),
                             0 if x_out is None else x_out.stride(-2), gamma.data_ptr(), beta.data_ptr(),
                             float(eps), out.data_ptr(), out.stride(-2), _f16(x), _stream()), "sa_layernorm")
    return out


# ----------------------------------------------------------------------------- small / memory-bound ops
@_launcher("embedding")
def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
                 silu_in: bool = False, silu_out: bool = False, accumulate: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 x [m, k] (row stride arbitrary), bf16 w [n, k] -> fp32 [m, n] (``accumulate``: out += ...)."""
    if x.dtype != torch.float32 or x.stride(-1) != 1:
        raise ValueError("x must be fp32 with unit inner stride")
    _chk_bf16(w, "w")
    if w.dtype != torch.bfloat16:
        raise ValueError("small_linear weights are bf16 in both activation modes")
    m, k = x.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=torch.float32)
    check(lib().sa_small_linear(x.data_ptr(), x.stride(0), m, w.data_ptr(), k, n, _ptr(bias), int(silu_in),
                                int(silu_out), int(accumulate), out.data_ptr(), out.stride(0), _stream()),
          "sa_small_linear")
    return out


@_launcher("embedding")
def sinusoid(t: torch.Tensor, dim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk_f32(t, "t")
    n = t.numel()
    if out is None:
        out = torch.empty((n, dim), device=t.device, dtype=torch.float32)
    check(lib().sa_sinusoid(t.data_ptr(), n, dim, out.data_ptr(), _stream()), "sa_sinusoid")
    return out


@_launcher("conv_direct")
def conv_direct(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor], cout: int, ksz: int,
                stride: int, pad: int, *, silu: bool = False, scale: float = 1.0,
                add: Optional[torch.Tensor] = None, add_from: int = 0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 [n, h, w, cin] (NHWC); wt bf16 [k*k*cin, cout_pad] -> bf16 [n, ho, wo, cout]."""
    _chk_bf16(x, "x")
    _chk_bf16(wt, "wt", x)
    n, h, w, cin = x.shape
    ho, wo = (h + 2 * pad - ksz) // stride + 1, (w + 2 * pad - ksz) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout), device=x.device, dtype=x.dtype)
    check(lib().sa_conv_direct(x.data_ptr(), n, h, w, cin, wt.data_ptr(), _ptr(bias), cout, wt.shape[1], ksz,
                               stride, pad, int(silu), float(scale), _ptr(add), add_from, out.data_ptr(),
                               _f16(x), _stream()), "sa_conv_direct")
    return out


@_launcher("layout")
def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk_bf16(x, "x")
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=x.dtype)
    check(lib().sa_upsample2x(x.data_ptr(), n, h, w, c, out.data_ptr(), _stream()), "sa_upsample2x")
    return out


@_launcher("layout")
def im2col_s2(x: torch.Tensor, out: Optional[torch.Tensor] = None, pad: int = 1) -> torch.Tensor:
    """3x3 stride-2 gather: pad=1 symmetric (UNet), pad=0 = F.pad(x, (0, 1, 0, 1)) (VAE encoder)."""
    _chk_bf16(x, "x")
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n * (h // 2) * (w // 2), 9 * c), device=x.device, dtype=x.dtype)
    if pad == 1:
        check(lib().sa_im2col_s2(x.data_ptr(), n, h, w, c, out.data_ptr(), _stream()), "sa_im2col_s2")
    else:
        check(lib().sa_im2col_s2_pad(x.data_ptr(), n, h, w, c, pad, out.data_ptr(), _stream()), "sa_im2col_s2_pad")
    return out


@_launcher("layout")
def im2col3x3_c8(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv_in gather: x bf16 [n, h, w, 8] -> bf16 [n*h*w, 128] (72 values (ky, kx, ci) + zero padding)."""
    _chk_bf16(x, "x")
    n, h, w, c = x.shape
    if c != 8 or not x.is_contiguous():
        raise ValueError("x must be contiguous [n, h, w, 8]")
    if out is None:
        out = torch.empty((n * h * w, 128), device=x.device, dtype=x.dtype)
    check(lib().sa_im2col3x3_c8(x.data_ptr(), n, h, w, out.data_ptr(), _stream()), "sa_im2col3x3_c8")
    return out


@_launcher("layout")
def nchw_to_nhwc(x: torch.Tensor, out: Optional[torch.Tensor] = None, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """[n, c, h, w] contiguous, fp32 or already in the activation format ``dtype`` -> ``dtype`` [n, h, w, c]."""
    if dtype not in ACT_DTYPES or not x.is_contiguous() or x.dtype not in (torch.float32, dtype):
        raise ValueError(f"x must be contiguous NCHW in fp32 or {dtype}")
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h, w, c), device=x.device, dtype=dtype)
    check(lib().sa_nchw_to_nhwc(x.data_ptr(), int(x.dtype != torch.float32), n, c, h * w, out.data_ptr(),
                                int(dtype == torch.float16), _stream()), "sa_nchw_to_nhwc")
    return out


@_launcher("layout")
def nhwc_to_nchw_f32(x: torch.Tensor, n: int, c: int, hw: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [n*hw, ld] (first c columns) -> fp32 [n, c, hw]."""
    if x.dtype != torch.float32:
        raise ValueError("x must be fp32")
    if out is None:
        out = torch.empty((n, c, hw), device=x.device, dtype=torch.float32)
    check(lib().sa_nhwc_to_nchw_f32(x.data_ptr(), x.stride(-2), n, c, hw, out.data_ptr(), _stream()),
          "sa_nhwc_to_nchw_f32")
    return out


@_launcher("sampler")
def sampler_prepare(latents, image_latents, idx, nb, b0, f, hw, sigma, out) -> None:
    check(lib().sa_sampler_prepare(latents.data_ptr(), image_latents.data_ptr(), idx.data_ptr(), nb, b0,
                                   idx.numel(), f, hw, float(sigma), out.data_ptr(), _f16(out), _stream()),
          "sa_sampler_prepare")


@_launcher("sampler")
def sampler_accumulate(unet_out, idx, weight, nb, b0, f, hw, noise) -> None:
    check(lib().sa_sampler_accumulate(unet_out.data_ptr(), unet_out.stride(-2), idx.data_ptr(),
                                      weight.data_ptr(), nb, b0, idx.numel(), f, hw, noise.data_ptr(),
                                      _stream()), "sa_sampler_accumulate")


@_launcher("sampler")
def sampler_step(noise, cnt, guidance, f, chw, sigma, sigma_next, latents) -> None:
    check(lib().sa_sampler_step(noise.data_ptr(), cnt.data_ptr(), guidance.data_ptr(), f, chw, float(sigma),
                                float(sigma_next), latents.data_ptr(), _stream()), "sa_sampler_step")
