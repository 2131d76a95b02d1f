"""One 3x3 convolution launch of the dominant shape (B=2 x 16 frames, 64x64, 320 -> 320, bias + per-batch row bias,
as ResnetBlock2D.conv1 runs it) for an `ncu --set full` capture:
  ncu --set full --clock-control none --import-source on -k regex:gemm_tap_persist -s 2 -c 1 -o out python tools/conv_once.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402

f, h, w, c = 32, 64, 64, 320
x = torch.randn(f, h, w, c, device="cuda").bfloat16()
wt = (torch.randn(c, 9 * c, device="cuda") * (9 * c) ** -0.5).bfloat16()
bias = torch.randn(c, device="cuda")
temb = torch.randn(2, c, device="cuda")
for _ in range(3):
    out = ops.gemm_tap(x, wt, grid=(w, h, f), taps=ops.TAPS_3X3, bias=bias, row_bias=temb, rows_per_group=16 * h * w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gemm_tap(x, wt, grid=(w, h, f), taps=ops.TAPS_3X3, bias=bias, row_bias=temb, rows_per_group=16 * h * w)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"conv3x3 320->320 @ 32x64x64: {ms * 1e3:.1f} us, {2.0 * f * h * w * c * 9 * c / ms / 1e9:.0f} TFLOP/s")
