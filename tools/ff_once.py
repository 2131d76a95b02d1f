"""Two sa_ff_fused launches at the 64x64 level of a B=2 call (M = 131072, C = 320), for ncu."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

m, c = 131072, 320
x = torch.randn(m, c, device="cuda").bfloat16()
w1 = (torch.randn(8 * c, c, device="cuda") * c ** -0.5).bfloat16()
b1 = torch.randn(8 * c, device="cuda")
w2 = (torch.randn(c, 4 * c, device="cuda") * (4 * c) ** -0.5).bfloat16()
b2 = torch.randn(c, device="cuda")
g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
addv = torch.randn(2, c, device="cuda")
wi, bi = ops.interleave_geglu(w1, b1)
out = torch.empty(m, c, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.ff_fused(x, g, b, wi, bi, w2, b2, addv=addv, rows_per_group=m // 2, out=out)
torch.cuda.synchronize()
print("ok")
