"""Time sa_gemm_tap on one shape: python tools/gemm_one.py M K N [geglu] [tile]  (CUDA events, warm L2)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402


def main():
    m, k, n = (int(x) for x in sys.argv[1:4])
    geglu = len(sys.argv) > 4 and sys.argv[4] == "1"
    tile = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
    b = torch.randn(n, device="cuda")
    res = None if geglu else torch.randn(m, n, device="cuda").bfloat16()
    for name, kw in (("bias", dict(bias=b)), ("bias+residual", dict(bias=b, residual=res))):
        if geglu and name != "bias":
            continue
        for _ in range(3):
            ops.gemm_tap(a, w, geglu=geglu, tile_n=tile, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm_tap(a, w, geglu=geglu, tile_n=tile, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"M={m} K={k} N={n} geglu={geglu} tile={tile or 'auto'} {name}: {ms * 1e3:.1f} us  {2.0 * m * k * n / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
