import sys, torch
sys.path.insert(0, "/root/repo")
from stableanimator_b200 import ops
dev = "cuda"
def bench(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, f, h, w_, ci, co in (("L2 conv 1280 B=2", 32, 16, 16, 1280, 1280), ("L2 conv 2560->1280 B=2", 32, 16, 16, 2560, 1280),
                                ("L1 conv 640 B=2", 32, 32, 32, 640, 640), ("L3 conv 1280 B=2", 32, 8, 8, 1280, 1280)):
    x = torch.randn(f, h, w_, ci, device=dev).bfloat16()
    wt = (torch.randn(co, 9 * ci, device=dev) * (9 * ci) ** -0.5).bfloat16()
    for t in (0, 64, 128, 160, 256):
        if t and co % t: continue
        ms = bench(lambda: ops.gemm_tap(x, wt, grid=(w_, h, f), taps=ops.TAPS_3X3, tile_n=t))
        tf = 2.0 * f * h * w_ * 9 * ci * co / ms / 1e9
        print(f"{name}: tile {t or 'auto'}: {ms*1e3:7.1f} us {tf:7.0f} TFLOP/s", flush=True)
