"""Fixed cost vs streaming rate of the streamed LayerNorm / GroupNorm kernels: time at several row counts (C = 320),
inputs AND outputs rotating over > 126 MB (cold reads, cold writes), and fit t = t0 + bytes / BW.  Also torch's device copy of
the same bytes for comparison."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402


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


def main():
    dev, c = "cuda", int(sys.argv[1]) if len(sys.argv) > 1 else 320
    g, b = torch.randn(c, device=dev), torch.randn(c, device=dev)
    res = {}
    for rows in (16384, 32768, 65536, 131072, 262144, 524288):
        nbuf = max(2, int(400e6 // (rows * c * 2)) + 1)
        xs = [torch.randn(rows, c, device=dev).bfloat16() for _ in range(nbuf)]
        ys = [torch.empty(rows, c, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        ws = torch.zeros(rows // 4096, 32, 2, device=dev, dtype=torch.float64)
        mb = 2.0 * rows * c * 2 / 1e6
        t_ln = timeit(lambda i: ops.layernorm(xs[i], g, b, out=ys[i]), nbuf)
        t_gn = timeit(lambda i: ops.groupnorm(xs[i], g, b, 1e-5, 4096, silu=True, stats_ws=ws, out=ys[i], stats_ready=True), nbuf)
        t_cp = timeit(lambda i: ys[i].copy_(xs[i]), nbuf)
        res[rows] = (mb, t_ln, t_gn, t_cp)
        print(f"rows {rows:7d} ({mb:6.1f} MB r+w): layernorm {t_ln:6.1f} us  groupnorm-apply {t_gn:6.1f} us  copy {t_cp:6.1f} us")
        del xs, ys
    (m0, *a), (m1, *bb) = res[131072], res[524288]
    for name, t0, t1 in zip(("layernorm", "groupnorm-apply", "copy"), a, bb):
        bw = (m1 - m0) / (t1 - t0)
        print(f"{name:16s}: streaming {bw:5.2f} TB/s, fixed cost {t0 - m0 / bw:5.1f} us per launch (in a graph)")


if __name__ == "__main__":
    main()
