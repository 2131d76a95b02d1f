"""A few GroupNorm / LayerNorm calls at UNet shapes, for `ncu --metrics gpu__time_duration.sum` or --set full."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

frames = 32
for s, c, rps in ((4096, 320, 4096), (4096, 320, 16 * 4096), (1024, 640, 1024), (4096, 960, 4096)):
    rows = frames * s
    x = torch.randn(rows, c, device="cuda").bfloat16()
    g, b = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
    for _ in range(2):
        y = ops.groupnorm(x, g, b, 1e-5, rps, silu=True)
    y2 = ops.layernorm(x, g, b) if c != 960 else None
torch.cuda.synchronize()
print("ok")
