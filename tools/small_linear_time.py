import sys, torch
sys.path.insert(0, "/root/repo")
from stableanimator_b200 import ops
x = torch.randn(2, 1280, device="cuda")
for n in (8832, 35000):
    w = (torch.randn(n, 1280, device="cuda") * 0.02).bfloat16()
    b = torch.randn(n, device="cuda")
    out = torch.empty(2, n, device="cuda")
    for silu in (True, False):
        for _ in range(3): ops.small_linear(x, w, b, silu_in=silu, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): ops.small_linear(x, w, b, silu_in=silu, out=out)
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"small_linear m=2 k=1280 n={n} silu_in={silu}: {us:.1f} us ({n*1280*2/us/1e3:.0f} GB/s, L2-resident weights if < 126 MB)")
