"""Python-side launchers: torch tensors in, raw pointers through the C ABI, torch tensors out.

PyTorch is used for device memory and streams only -- every arithmetic op below runs in
``libsa_b200.so``.  Layout convention: activations are bf16 and channel-innermost
(``[frames, h, w, C]`` == ``[frames*h*w, C]``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import GemmArgs, check, lib

TAPS_1 = ((0, 0, 0),)
# 3x3 conv, pad 1: tap index = ky*3 + kx, offset (dx, dy, 0) = (kx-1, ky-1, 0)
TAPS_3X3 = tuple((kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3))
# temporal conv k=(3,1,1), pad (1,0,0) on a (S, F, B) grid: offset along dim 2 (frames)
TAPS_T3 = ((0, -1, 0), (0, 0, 0), (0, 1, 0))


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk_bf16(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA bf16 tensor, got {t.dtype} on {t.device}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must be channel-innermost (stride(-1) == 1)")


def _chk_f32(t: Optional[torch.Tensor], name: str) -> None:
    if t is None:
        return
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA fp32 tensor")


def pick_tile(m: int, n: int, geglu: bool = False) -> int:
    return lib().sa_gemm_pick_tile(m, n, int(geglu))


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
) -> torch.Tensor:
    """``out = epilogue(sum_taps A(shifted) @ W^T)`` -- see ``sa_gemm_tap`` in include/sa_b200.h.

    ``a0`` / ``a1``: bf16, ``[..., C]`` with a uniform row stride (``stride(-2)``), all leading dims
    dense over the position grid ``grid = (d1, d2, d3)`` (default ``(rows, 1, 1)``).
    ``w``: bf16 ``[N, ntaps*(C0+C1)]``.
    """
    _chk_bf16(a0, "a0")
    _chk_bf16(w, "w")
    c0 = a0.shape[-1]
    lda0 = a0.stride(-2) if a0.dim() >= 2 else c0
    rows = a0.numel() // c0
    c1, lda1 = 0, 0
    if a1 is not None:
        _chk_bf16(a1, "a1")
        c1 = a1.shape[-1]
        lda1 = a1.stride(-2)
        if a1.numel() // c1 != rows:
            raise ValueError("a0 and a1 must cover the same positions")
    if grid is None:
        grid = (rows, 1, 1)
    d1, d2, d3 = (int(x) for x in grid)
    if d1 * d2 * d3 != rows:
        raise ValueError(f"grid {grid} does not match {rows} positions")
    ntaps = len(taps)
    n_mma = int(n) if n is not None else w.shape[0]
    if w.shape[1] != ntaps * (c0 + c1):
        raise ValueError(f"w has K={w.shape[1]}, expected {ntaps}*({c0}+{c1})")
    n_out = n_mma // 2 if geglu else n_mma
    if out is None:
        out = torch.empty((rows, n_out), device=a0.device,
                          dtype=torch.float32 if out_f32 else torch.bfloat16)
    _chk_f32(bias, "bias")
    _chk_f32(row_bias, "row_bias")
    _chk_f32(row_scale, "row_scale")
    args = GemmArgs()
    args.a0, args.a1 = a0.data_ptr(), _ptr(a1)
    args.c0, args.c1, args.lda0, args.lda1 = c0, c1, lda0, lda1
    args.d1, args.d2, args.d3 = d1, d2, d3
    args.ntaps = ntaps
    for i, t in enumerate(taps):
        args.tap_off[3 * i], args.tap_off[3 * i + 1], args.tap_off[3 * i + 2] = t
    args.w, args.n = w.data_ptr(), n_mma
    args.bias, args.row_bias, args.row_scale = _ptr(bias), _ptr(row_bias), _ptr(row_scale)
    args.rows_per_group = rows_per_group
    if residual is not None:
        _chk_bf16(residual, "residual")
        args.residual, args.ldr = residual.data_ptr(), residual.stride(-2)
    if blend_x is not None:
        _chk_bf16(blend_x, "blend_x")
        args.blend_x, args.ldx = blend_x.data_ptr(), blend_x.stride(-2)
    args.alpha = float(alpha)
    args.geglu = int(geglu)
    args.out, args.ldc = out.data_ptr(), out.stride(-2)
    args.out_f32 = int(out.dtype == torch.float32)
    args.tile_n = tile_n
    check(lib().sa_gemm_tap(C.byref(args), _stream()), "sa_gemm_tap")
    return out


def interleave_geglu(w: torch.Tensor, b: Optional[torch.Tensor]):
    """Reorder GEGLU projection rows ``[value(4C) ; gate(4C)]`` (unet.py:865-870) into 64-wide
    blocks ``[v0..63 | g0..63 | v64..127 | g64..127 | ...]`` so one MMA tile holds both halves."""
    n2 = w.shape[0]
    h = n2 // 2
    assert h % 64 == 0
    idx = torch.arange(n2, device=w.device).reshape(2, h // 64, 64).permute(1, 0, 2).reshape(-1)
    return w[idx].contiguous(), (None if b is None else b[idx].contiguous())
