"""Bandwidth of the memory-bound kernels at the UNet's shapes (B=2 call): GroupNorm(+SiLU), LayerNorm,
ID attention.  Inputs rotate over enough buffers to exceed the 126 MB L2 (cold reads, like the ncu launch
list); algorithmic bytes = read x once + write y once."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402


def timeit(fn, nbuf, iters=12):
    """Device time per launch, replayed from a CUDA graph (eager Python launches of 30-60 us kernels are host-bound)."""
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
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    dev = "cuda"
    frames = 32
    for name, s, c in (("L0", 4096, 320), ("L0 cat", 4096, 640), ("L1", 1024, 640), ("L2", 256, 1280), ("L3", 64, 1280)):
        rows = frames * s
        nbuf = max(2, int(300e6 // (rows * c * 2)) + 1)
        xs = [torch.randn(rows, c, device=dev).bfloat16() for _ in range(nbuf)]
        y = torch.empty(rows, c, device=dev, dtype=torch.bfloat16)
        g, b = torch.randn(c, device=dev), torch.randn(c, device=dev)
        byt = 2.0 * rows * c * 2
        for label, rps in (("per-frame", s), ("per-tile", 16 * s)):
            ws = torch.empty(rows // rps, 32, 2, device=dev, dtype=torch.float64)
            us = timeit(lambda i: ops.groupnorm(xs[i], g, b, 1e-5, rps, silu=True, stats_ws=ws, out=y), nbuf)
            print(f"groupnorm+silu {name:7s} C={c:4d} {label:9s}: {us:7.1f} us  {byt / us / 1e3:7.0f} GB/s (algorithmic)")
        us = timeit(lambda i: ops.layernorm(xs[i], g, b, out=y), nbuf)
        print(f"layernorm      {name:7s} C={c:4d}          : {us:7.1f} us  {byt / us / 1e3:7.0f} GB/s (algorithmic)")
        if name in ("L0", "L1", "L2"):
            heads = c // 64
            ik = torch.randn(2, 4, c, device=dev)
            iv = torch.randn(2, 4, c, device=dev)
            st = torch.empty(frames, 2, device=dev, dtype=torch.float64)
            us = timeit(lambda i: ops.attn_id(xs[i], ik, iv, 2, 16, s, heads, st, out=y), nbuf)
            print(f"attn_id        {name:7s} C={c:4d}          : {us:7.1f} us  {byt / us / 1e3:7.0f} GB/s (algorithmic)")
        del xs


if __name__ == "__main__":
    main()
