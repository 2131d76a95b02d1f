"""Per-launch cost inside a CUDA graph with and without programmatic dependent launch (SA_PDL, read per launch):
a chain of N small LayerNorm launches and a chain of N small GEMMs, captured twice, replayed alternately."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

N = 300
dev = "cuda"
x = torch.randn(4096, 320, device=dev).bfloat16()
g_, b_ = torch.ones(320, device=dev), torch.zeros(320, device=dev)
bufs = [torch.empty_like(x) for _ in range(2)]
a = torch.randn(148 * 128, 320, device=dev).bfloat16()
w = torch.randn(320, 320, device=dev).bfloat16() * 0.05
gb = [torch.empty_like(a) for _ in range(2)]


def chain_ln():
    src = x
    for i in range(N):
        ops.layernorm(src, g_, b_, out=bufs[i & 1])
        src = bufs[i & 1]


def chain_gemm():
    src = a
    for i in range(N):
        ops.gemm_tap(src, w, out=gb[i & 1])
        src = gb[i & 1]


def chain_mixed():
    src = a
    for i in range(N // 2):
        ops.layernorm(src, g_[:320], b_[:320], out=gb[0])
        ops.gemm_tap(gb[0], w, out=gb[1])
        src = gb[1]


for name, fn in (("layernorm 4096x320", chain_ln), ("gemm 18944x320x320", chain_gemm),
                 ("layernorm -> gemm 18944x320x320", chain_mixed)):
    graphs = {}
    for val in ("0", "1", "2"):
        os.environ["SA_PDL"] = val
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        graphs[val] = g
    res = {k: [] for k in graphs}
    for _ in range(10):
        for val, g in graphs.items():
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            res[val].append(e0.elapsed_time(e1) * 1e3 / N)
    for val, ts in res.items():
        ts = sorted(ts)
        print(f"{name}: SA_PDL={val}: {ts[len(ts) // 2]:.2f} us per launch (min {ts[0]:.2f})")
