"""Diagnostics for sa_attn_spatial: error statistics vs fp64 SDPA and timing."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from stableanimator_b200 import ops  # noqa: E402
from util import excess, rel_l2  # noqa: E402


def ref_attn(qkv, frames, s, heads):
    c = heads * 64
    q, k, v = (t.float().reshape(frames, s, heads, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    o = torch.nn.functional.scaled_dot_product_attention(q.double(), k.double(), v.double())
    return o.permute(0, 2, 1, 3).reshape(frames * s, c).float()


def main():
    for frames, s, heads, scale in [(3, 1024, 1, 1.0), (2, 256, 2, 1.0), (1, 576, 2, 1.0), (2, 4096, 5, 1.0), (2, 1024, 2, 3.0),
                                    (1, 2048, 1, 6.0)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        qkv = (torch.randn(frames * s, 3 * heads * 64, generator=g, device="cuda") * scale).bfloat16()
        ref = ref_attn(qkv, frames, s, heads)
        out = ops.attn_spatial(ops.v_block_to_f16_(qkv, heads * 64), frames, s, heads)
        err = (out.float() - ref).abs()
        i = err.argmax().item()
        r, cc = divmod(i, ref.shape[1])
        print(f"F={frames} S={s} H={heads} scale={scale}: excess={excess(out, ref):.3e} rel_l2={rel_l2(out, ref):.3e} "
              f"max|ref|={ref.abs().max().item():.3f} max err {err.max().item():.3e} at row {r} col {cc} "
              f"(ref {ref[r, cc].item():.4f} out {out[r, cc].item():.4f}); rows with err>1e-3*max: "
              f"{(err.max(dim=1).values > 1e-3 * ref.abs().max()).sum().item()}")
    # timing at the L0 shape (B=2: 32 frames, S=4096, 5 heads) and L1
    for frames, s, heads in [(32, 4096, 5), (32, 1024, 10), (32, 256, 20)]:
        qkv = torch.randn(frames * s, 3 * heads * 64, device="cuda").bfloat16()
        out = torch.empty(frames * s, heads * 64, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.attn_spatial(qkv, frames, s, heads, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attn_spatial(qkv, frames, s, heads, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * frames * heads * s * s * 64
        print(f"time F={frames} S={s} H={heads}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
