"""sa_ff_fused against the kernel chain it replaces (LayerNorm -> GEGLU GEMM -> out-projection GEMM) at the 64x64 level
of a B=2 call (M = 131072, C = 320) and at 128x72 (M = 294912): CUDA events, inputs rotating over > 126 MB."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402


def timeit(fn, nbuf, iters=10):
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for m, c in ((131072, 320), (294912, 320), (131072, 256)):
    nbuf = 3
    xs = [torch.randn(m, c, device="cuda").bfloat16() for _ in range(nbuf)]
    w1 = (torch.randn(8 * c, c, device="cuda") * c ** -0.5).bfloat16()
    b1 = torch.randn(8 * c, device="cuda")
    w2 = (torch.randn(c, 4 * c, device="cuda") * (4 * c) ** -0.5).bfloat16()
    b2 = torch.randn(c, device="cuda")
    g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    addv = torch.randn(2, c, device="cuda")
    wi, bi = ops.interleave_geglu(w1, b1)
    out = torch.empty(m, c, device="cuda", dtype=torch.bfloat16)
    z0 = torch.empty(m, c, device="cuda", dtype=torch.bfloat16)
    a = torch.empty(m, c, device="cuda", dtype=torch.bfloat16)
    hid = torch.empty(m, 4 * c, device="cuda", dtype=torch.bfloat16)

    def chain(i):
        ops.layernorm(xs[i], g, b, addv=addv, rows_per_group=m // 2, x_out=z0, out=a)
        ops.gemm_tap(a, wi, bias=bi, geglu=True, out=hid)
        ops.gemm_tap(hid, w2, bias=b2, residual=z0, out=out)

    def fused(i):
        ops.ff_fused(xs[i], g, b, wi, bi, w2, b2, addv=addv, rows_per_group=m // 2, out=out)

    fl = 2.0 * m * c * 8 * c + 2.0 * m * 4 * c * c
    tc, tf = timeit(chain, nbuf), timeit(fused, nbuf)
    print(f"M={m} C={c}: chain {tc:7.1f} us ({fl / tc / 1e6:6.0f} TFLOP/s)   fused {tf:7.1f} us ({fl / tf / 1e6:6.0f} TFLOP/s)")
