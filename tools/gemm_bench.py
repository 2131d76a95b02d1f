"""Micro-benchmark of sa_gemm_tap on the UNet's dominant shapes (CUDA events, L2-cold by rotation)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402


def bench(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = []
    dev = "cuda"
    shapes = [  # (name, M, K, N, geglu)
        ("L0 qkv", 65536, 320, 960, False), ("L0 proj", 65536, 320, 320, False),
        ("L0 geglu", 65536, 320, 2560, True), ("L0 ffout", 65536, 1280, 320, False),
        ("L1 qkv", 16384, 640, 1920, False), ("L1 geglu", 16384, 640, 5120, True),
        ("L1 ffout", 16384, 2560, 640, False),
        ("L2 qkv", 4096, 1280, 3840, False), ("L2 geglu", 4096, 1280, 10240, True),
        ("L2 ffout", 4096, 5120, 1280, False), ("L3 proj", 1024, 1280, 1280, False),
    ]
    tiles = [0, 64, 128, 160, 256]
    for name, m, k, n, geglu in shapes:
        a = torch.randn(m, k, device=dev).bfloat16()
        w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
        b = torch.randn(n, device=dev)
        for t in tiles:
            if t and (n % t or (geglu and t not in (128, 256))):
                continue
            try:
                ms = bench(lambda: ops.gemm_tap(a, w, bias=b, geglu=geglu, tile_n=t))
            except Exception as e:  # noqa: BLE001
                out.append({"name": name, "tile": t, "error": str(e)[:200]})
                continue
            tf = 2.0 * m * k * n / ms / 1e9
            out.append({"name": name, "tile": t or ops.pick_tile(m, n, geglu), "auto": t == 0,
                        "ms": round(ms, 4), "tflops": round(tf, 1)})
            print(out[-1], flush=True)
        ms = bench(lambda: torch.nn.functional.linear(a, w))
        out.append({"name": name, "tile": "cublas", "ms": round(ms, 4),
                    "tflops": round(2.0 * m * k * n / ms / 1e9, 1)})
        print(out[-1], flush=True)
    # 3x3 convs (implicit GEMM)
    convs = [("L0 conv 320", 16, 64, 64, 320, 320), ("L1 conv 640", 16, 32, 32, 640, 640),
             ("L2 conv 1280", 16, 16, 16, 1280, 1280), ("L3 conv 1280", 16, 8, 8, 1280, 1280),
             ("L2 conv 2560->1280", 16, 16, 16, 2560, 1280)]
    for name, f, h, w_, ci, co in convs:
        x = torch.randn(f, h, w_, ci, device=dev).bfloat16()
        wt = (torch.randn(co, 9 * ci, device=dev) * (9 * ci) ** -0.5).bfloat16()
        for t in tiles:
            if t and co % t:
                continue
            ms = bench(lambda: ops.gemm_tap(x, wt, grid=(w_, h, f), taps=ops.TAPS_3X3, tile_n=t))
            tf = 2.0 * f * h * w_ * 9 * ci * co / ms / 1e9
            out.append({"name": name, "tile": t or ops.pick_tile(f * h * w_, co, False),
                        "auto": t == 0, "ms": round(ms, 4), "tflops": round(tf, 1)})
            print(out[-1], flush=True)
        xn = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wn = wt.reshape(co, 3, 3, ci).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        ms = bench(lambda: torch.nn.functional.conv2d(xn, wn, padding=1))
        out.append({"name": name, "tile": "cudnn", "ms": round(ms, 4),
                    "tflops": round(2.0 * f * h * w_ * 9 * ci * co / ms / 1e9, 1)})
        print(out[-1], flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/gemm_bench.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
