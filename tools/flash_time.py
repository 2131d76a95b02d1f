"""Timing of sa_attn_spatial at the UNet's shapes (CUDA events, 20 launches): python tools/flash_time.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

for frames, s, heads in [(32, 4096, 5), (32, 1024, 10), (32, 256, 20), (32, 9216, 5)]:
    qkv = torch.randn(frames * s, 3 * heads * 64, device="cuda").bfloat16()
    ops.v_block_to_f16_(qkv, heads * 64)
    out = torch.empty(frames * s, heads * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attn_spatial(qkv, frames, s, heads, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attn_spatial(qkv, frames, s, heads, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 4.0 * frames * heads * s * s * 64
    print(f"time F={frames} S={s} H={heads}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
