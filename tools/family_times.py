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


from stableanimator_b200.profiling import MEMBERS, elided, graph_ms  # noqa: E402


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
        members = MEMBERS.get(fam, (fam,))
        origs_m = {m: getattr(ops, m) for m in members}
        for m in members:
            setattr(ops, m, elided(m))
        try:
            rest = graph_ms(run, a.reps)
        finally:
            for m in members:
                setattr(ops, m, origs_m[m])
        res[fam] = full - rest
        tot += full - rest
        print(f"  {fam:16s} {full - rest:8.3f} ms  {100 * (full - rest) / full:5.1f}%")
    print(f"  {'(other)':16s} {full - tot:8.3f} ms  {100 * (full - tot) / full:5.1f}%")
    # every family elided at once: what is left is the small kernels outside the families plus launch gaps
    fams = ("gemm_tap", "ff_fused", "groupnorm", "layernorm", "attn_spatial", "attn_temporal", "attn_id")
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
