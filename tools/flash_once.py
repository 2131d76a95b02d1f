"""One sa_attn_spatial launch at the L0 shape of a B=2 call (32 frames, S=4096, 5 heads) for ncu."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

frames, s, heads = 32, 4096, 5
qkv = torch.randn(frames * s, 3 * heads * 64, device="cuda").bfloat16()
ops.v_block_to_f16_(qkv, heads * 64)
out = torch.empty(frames * s, heads * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_spatial(qkv, frames, s, heads, out=out)
torch.cuda.synchronize()
print("ok")
