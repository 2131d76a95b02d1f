"""One GEGLU GEMM launch set (L0 shape of a B=2 call) for ncu."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

m, k, n = 131072, 320, 2560
tile = int(sys.argv[1]) if len(sys.argv) > 1 else 0
a = torch.randn(m, k, device="cuda").bfloat16()
w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
b = torch.randn(n, device="cuda")
for _ in range(3):
    ops.gemm_tap(a, w, bias=b, geglu=True, tile_n=tile)
torch.cuda.synchronize()
print("ok")
