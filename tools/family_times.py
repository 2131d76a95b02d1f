"""In-situ device time of every kernel family inside one UNet call (CUDA events, graph replay).

For each family the same call is captured with that family's launches elided (its outputs stay
uninitialised; no kernel has data-dependent control flow), and the difference to the full graph is
the family's time with the real cache state and launch overlap -- unlike an ncu launch list, which
is serialised and cold-cache.  Prints one line per family plus the sum.

  python tools/family_times.py [--latent 64] [--frames 16] [--batch 2] [--size full]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import ops  # noqa: E402
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict  # noqa: E402
from stableanimator_b200.unet import UNetB200  # noqa: E402


def _empty_like_out(rows, cols, dev, dtype=torch.bfloat16):
    return torch.empty((rows, cols), device=dev, dtype=dtype)


def elided(name):
    if name == "gemm_tap":
        def f(a0, w_, **kw):
            rows = a0.numel() // a0.shape[-1] // (kw.get("in_stride", 1) ** 2)   # output positions
            n = kw.get("n") or w_.shape[0]
            if kw.get("out") is not None:
                return kw["out"]
            return _empty_like_out(rows, n // 2 if kw.get("geglu") else n, a0.device,
                                   torch.float32 if kw.get("out_f32") else torch.bfloat16)
        return f
    if name == "groupnorm":
        def f(x0, gamma, beta, eps, rps, *, x1=None, silu=True, stats_ws=None, out=None, stats_ready=False):
            c = x0.shape[-1] + (0 if x1 is None else x1.shape[-1])
            return out if out is not None else _empty_like_out(x0.numel() // x0.shape[-1], c, x0.device)
        return f
    if name == "layernorm":
        def f(x, gamma, beta, *, addv=None, rows_per_group=1, x_out=None, out=None, eps=1e-5):
            return out if out is not None else torch.empty_like(x)
        return f
    if name == "attn_spatial":
        return lambda qkv, frames, s, heads, out=None: _empty_like_out(frames * s, heads * 64, qkv.device)
    if name == "attn_temporal":
        return lambda qkv, b, f, s, heads, out=None: _empty_like_out(b * f * s, heads * 64, qkv.device)
    if name == "attn_id":
        return lambda q, ik, iv, b, f, s, heads, stats, out=None: _empty_like_out(b * f * s, heads * 64, q.device)
    raise KeyError(name)


def graph_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="full")
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    cfg = UNetConfig.full() if a.size == "full" else UNetConfig.tiny()
    unet = UNetB200(cfg, synth_unet_state_dict(cfg), "cuda")
    h, w = a.latent, (a.width or a.latent)
    inp = synth_inputs(cfg, a.frames, h, w)
    B, F = a.batch, a.frames
    ctx = unet.prepare_context(inp["encoder_hidden_states"][:B].cuda())
    ati = inp["added_time_ids"][:B].cuda()
    x = torch.randn(B * F, h, w, 8, device="cuda").bfloat16()
    t = torch.tensor(1.0, device="cuda")
    run = lambda: unet.forward_nhwc(x, t, ctx, ati, B, F, h, w)  # noqa: E731
    full = graph_ms(run, a.reps)
    res = {"full_ms": full}
    print(f"UNet call B={B} F={F} {h}x{w}: {full:.3f} ms (CUDA graph replay)")
    tot = 0.0
    for fam in ("gemm_tap", "groupnorm", "layernorm", "attn_spatial", "attn_temporal", "attn_id"):
        orig = getattr(ops, fam)
        setattr(ops, fam, elided(fam))
        try:
            rest = graph_ms(run, a.reps)
        finally:
            setattr(ops, fam, orig)
        res[fam] = full - rest
        tot += full - rest
        print(f"  {fam:16s} {full - rest:8.3f} ms  {100 * (full - rest) / full:5.1f}%")
    print(f"  {'(other)':16s} {full - tot:8.3f} ms  {100 * (full - tot) / full:5.1f}%")
    # every family elided at once: what is left is the small kernels outside the families plus launch gaps
    fams = ("gemm_tap", "groupnorm", "layernorm", "attn_spatial", "attn_temporal", "attn_id")
    origs = {f: getattr(ops, f) for f in fams}
    for f in fams:
        setattr(ops, f, elided(f))
    try:
        res["none_ms"] = graph_ms(run, a.reps)
    finally:
        for f in fams:
            setattr(ops, f, origs[f])
    print(f"  all families elided: {res['none_ms']:.3f} ms remain (kernels outside the families)")
    if a.json:
        Path(a.json).write_text(json.dumps(res))


if __name__ == "__main__":
    main()
