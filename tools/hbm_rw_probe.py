"""Pure-write and pure-read bandwidth of the HBM at the sizes of the UNet's activations (torch fill / sum kernels; buffers rotate
over > 126 MB so nothing is L2-resident): calibrates what a write-heavy epilogue or a read-only statistics pass can reach.
(The read + write ceiling is the copy line of tools/norm_fit.py.)"""
import torch


def timeit(fn, nbuf, iters=20):
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


for mb in (21, 42, 84, 168, 336, 672):
    n = mb * 1000 * 1000 // 2
    nbuf = max(3, 600 // mb + 1)
    xs = [torch.randn(n, device="cuda").bfloat16() for _ in range(nbuf)]
    ys = [torch.empty(n, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    acc = torch.zeros((), device="cuda", dtype=torch.float32)
    t_w = timeit(lambda i: ys[i].fill_(1.0), nbuf)
    t_r = timeit(lambda i: torch.sum(xs[i], (0,), dtype=torch.float32, out=acc), nbuf)
    print(f"{mb:4d} MB: fill {t_w:6.1f} us ({mb / t_w:5.2f} TB/s)   sum {t_r:6.1f} us ({mb / t_r:5.2f} TB/s)")
