"""A/B of one UNet option inside ONE process: two captured graphs of the same B=2 call replayed alternately
(same clocks / thermal state), e.g. fused GroupNorm statistics on / off.  python tools/unet_ab.py [--latent 64]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict  # noqa: E402
from stableanimator_b200.unet import UNetB200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=8)
    a = ap.parse_args()
    cfg = UNetConfig.full()
    unet = UNetB200(cfg, synth_unet_state_dict(cfg), "cuda")
    h, w = a.latent, (a.width or a.latent)
    B, F = 2, 16
    inp = synth_inputs(cfg, F, h, w)
    ctx = unet.prepare_context(inp["encoder_hidden_states"][:B].cuda())
    ati = inp["added_time_ids"][:B].cuda()
    x = torch.randn(B * F, h, w, 8, device="cuda").bfloat16()
    t = torch.tensor(1.0, device="cuda")
    graphs = {}
    for name, fuse in (("fused", True), ("two-pass", False)):
        unet.fuse_gn_stats = fuse
        unet.forward_nhwc(x, t, ctx, ati, B, F, h, w)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            unet.forward_nhwc(x, t, ctx, ati, B, F, h, w)
        graphs[name] = g
    times = {k: [] for k in graphs}
    for _ in range(a.rounds):
        for name, g in graphs.items():
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) / 5)
    for name, ts in times.items():
        ts = sorted(ts)
        print(f"{name:9s}: median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f}  max {ts[-1]:.3f}")


if __name__ == "__main__":
    main()
