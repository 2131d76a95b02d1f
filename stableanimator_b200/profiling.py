"""Measurement helpers shared by ``bench.py`` and ``tools/`` (not on the product path).

* ``graph_ms`` -- CUDA-event time of a CUDA-graph replay of a callable.
* ``family_times`` -- in-situ device time of a kernel family inside one UNet call: the same call is captured with the
  family's launches elided (their outputs stay uninitialised; no kernel of the path has data-dependent control flow)
  and the difference to the full graph is the family's time with the step's real cache state and launch overlap.
* ``memory_bound_gbs`` -- the memory-bound kernels at the UNet's dominant shape, inputs rotating over more buffers than
  the 126 MB L2 holds, as algorithmic GB/s (read x once + write y once).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable

import torch

from . import ops

FAMILIES = ("gemm_tap", "groupnorm", "layernorm", "attn_spatial", "attn_temporal", "attn_id")
# launchers elided together under one family name: the tensor-core GEMM family is the tapped GEMM plus the fused
# feed-forward (two chained GEMMs in one kernel)
MEMBERS = {"gemm_tap": ("gemm_tap", "ff_fused")}


def _empty(rows, cols, dev, dtype=torch.bfloat16):
    return torch.empty((rows, cols), device=dev, dtype=dtype)


def elided(name: str, on_call: Callable = None):
    """A stand-in for ``ops.<name>`` that launches nothing and returns an output buffer of the right shape."""
    if name == "gemm_tap":
        def f(a0, w_, **kw):
            rows = a0.numel() // a0.shape[-1] // (kw.get("in_stride", 1) ** 2)   # output positions
            n = kw.get("n") or w_.shape[0]
            if on_call is not None:
                on_call(2.0 * rows * w_.shape[1] * n)
            if kw.get("out") is not None:
                return kw["out"]
            return _empty(rows, n // 2 if kw.get("geglu") else n, a0.device,
                          torch.float32 if kw.get("out_f32") else torch.bfloat16)
        return f
    if name == "ff_fused":
        def f(x, gamma, beta, w1, b1, w2, b2, **kw):
            c = x.shape[-1]
            rows = x.numel() // c
            if on_call is not None:
                on_call(2.0 * rows * c * 8 * c + 2.0 * rows * 4 * c * c)
            return kw["out"] if kw.get("out") is not None else _empty(rows, c, x.device)
        return f
    if name == "groupnorm":
        def f(x0, gamma, beta, eps, rps, *, x1=None, silu=True, stats_ws=None, out=None, stats_ready=False):
            c = x0.shape[-1] + (0 if x1 is None else x1.shape[-1])
            return out if out is not None else _empty(x0.numel() // x0.shape[-1], c, x0.device)
        return f
    if name == "layernorm":
        def f(x, gamma, beta, *, addv=None, rows_per_group=1, x_out=None, out=None, eps=1e-5):
            return out if out is not None else torch.empty_like(x)
        return f
    if name == "attn_spatial":
        return lambda qkv, frames, s, heads, out=None: _empty(frames * s, heads * 64, qkv.device)
    if name == "attn_temporal":
        return lambda qkv, b, f, s, heads, out=None: _empty(b * f * s, heads * 64, qkv.device)
    if name == "attn_id":
        return lambda q, ik, iv, b, f, s, heads, stats, out=None: _empty(b * f * s, heads * 64, q.device)
    raise KeyError(name)


def graph_ms(fn: Callable, reps: int = 5) -> float:
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def family_times(run: Callable, families: Iterable[str] = FAMILIES, reps: int = 5, flops: list = None) -> Dict[str, float]:
    """``run()`` = one UNet call through ``ops``.  Returns {"full_ms", <family>: ms, ...}; ``flops`` (a list) receives
    the algorithmic flops of every elided GEMM launch (two passes: warm-up + capture)."""
    res = {"full_ms": graph_ms(run, reps)}
    for fam in families:
        members = MEMBERS.get(fam, (fam,))
        origs = {m: getattr(ops, m) for m in members}
        for m in members:
            setattr(ops, m, elided(m, flops.append if (fam == "gemm_tap" and flops is not None) else None))
        try:
            res[fam] = res["full_ms"] - graph_ms(run, reps)
        finally:
            for m in members:
                setattr(ops, m, origs[m])
    return res


def _rotating_us(fn: Callable, nbuf: int, iters: int = 12) -> float:
    """Device time per launch of ``fn(i)`` over rotating inputs, replayed from a CUDA graph: launched eagerly from Python these
    30-50 us kernels are bound by the host's launch rate, not by the device."""
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def memory_bound_gbs(dev, frames: int = 32, s: int = 4096, c: int = 320, hbm_gbs: float = 6527.5) -> dict:
    """GroupNorm+SiLU (two-pass and apply-only with the statistics from the producing GEMM), LayerNorm, the
    fused CFG + Euler step and the time-embedding chain, at the L0 shape of a B=2 call."""
    rows = frames * s
    nbuf = max(2, int(300e6 // (rows * c * 2)) + 1)
    xs = [torch.randn(rows, c, device=dev).bfloat16() for _ in range(nbuf)]
    ys = [torch.empty(rows, c, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]   # outputs rotate too: cold writes
    g, b = torch.randn(c, device=dev), torch.randn(c, device=dev)
    byt = 2.0 * rows * c * 2
    out = {}

    def rec(name, us, nbytes):
        out[name] = {"us": round(us, 2), "GB/s": round(nbytes / us / 1e3, 1), "frac": round(nbytes / us / 1e3 / hbm_gbs, 4),
                     "bytes": int(nbytes)}

    ws = torch.zeros(frames, 32, 2, device=dev, dtype=torch.float64)
    rec("groupnorm_silu_two_pass", _rotating_us(lambda i: ops.groupnorm(xs[i], g, b, 1e-5, s, silu=True, stats_ws=ws, out=ys[i]), nbuf), byt)
    ops.groupnorm(xs[0], g, b, 1e-5, s, silu=True, stats_ws=ws, out=ys[0])   # valid statistics for the apply-only pass
    rec("groupnorm_silu_apply", _rotating_us(lambda i: ops.groupnorm(xs[i], g, b, 1e-5, s, silu=True, stats_ws=ws, out=ys[i],
                                                                      stats_ready=True), nbuf), byt)
    rec("layernorm", _rotating_us(lambda i: ops.layernorm(xs[i], g, b, out=ys[i]), nbuf), byt)
    del xs, ys
    # fused normalise + CFG + Euler step (sampler_step): 2 reads eps + 1 read x + 1 write x, fp32 [F, 4, h, w]
    f_, hw = 16, s
    noise = torch.randn(2, f_, 4, hw, device=dev)
    cnt = torch.ones(f_, device=dev)
    gd = torch.full((f_,), 3.0, device=dev)
    lat = torch.randn(f_, 4, hw, device=dev)
    rec("sampler_cfg_euler_step", _rotating_us(lambda i: ops.sampler_step(noise, cnt, gd, f_, 4 * hw, 1.0, 0.9, lat), 1, 50),
        4.0 * f_ * 4 * hw * 4)
    return out
