"""Run a few representative sa_gemm_tap launches between cudaProfilerStart/Stop (for ncu --set full)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

dev = "cuda"
cases = [("L0 qkv", 65536, 320, 960, False), ("L0 geglu", 65536, 320, 2560, True),
         ("L1 ffout", 16384, 2560, 640, False)]
ts = []
for name, m, k, n, geglu in cases:
    a = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
    b = torch.randn(n, device=dev)
    res = None if geglu else torch.randn(m, n, device=dev).bfloat16()
    ts.append((a, w, b, geglu, res))
x = torch.randn(16, 64, 64, 320, device=dev).bfloat16()
wc = (torch.randn(320, 9 * 320, device=dev) * (9 * 320) ** -0.5).bfloat16()
for _ in range(2):
    for a, w, b, geglu, res in ts:
        ops.gemm_tap(a, w, bias=b, geglu=geglu, residual=res)
    ops.gemm_tap(x, wc, grid=(64, 64, 16), taps=ops.TAPS_3X3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for a, w, b, geglu, res in ts:
    ops.gemm_tap(a, w, bias=b, geglu=geglu, residual=res)
ops.gemm_tap(x, wc, grid=(64, 64, 16), taps=ops.TAPS_3X3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
