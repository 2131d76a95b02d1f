"""One B=2 UNet call (config 2 shape by default) between cudaProfilerStart/Stop, for ncu launch lists:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/unet_once.py
Without ncu it prints the eager and CUDA-graph time of the call."""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict  # noqa: E402
from stableanimator_b200.unet import UNetB200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="full")
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    cfg = UNetConfig.full() if a.size == "full" else UNetConfig.tiny()
    unet = UNetB200(cfg, synth_unet_state_dict(cfg), "cuda")
    inp = synth_inputs(cfg, a.frames, a.latent, a.latent)
    B, F, h = a.batch, a.frames, a.latent
    ctx = unet.prepare_context(inp["encoder_hidden_states"][:B].cuda())
    ati = inp["added_time_ids"][:B].cuda()
    x = torch.randn(B * F, h, h, 8, device="cuda").bfloat16()
    t = torch.tensor(1.0, device="cuda")
    for _ in range(2):
        unet.forward_nhwc(x, t, ctx, ati, B, F, h, h)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    unet.forward_nhwc(x, t, ctx, ati, B, F, h, h)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"eager UNet call B={B} F={F} {h}x{h}: {(time.perf_counter() - t0) * 1e3:.2f} ms")
    if a.graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            unet.forward_nhwc(x, t, ctx, ati, B, F, h, h)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"graph UNet call: {e0.elapsed_time(e1) / 5:.2f} ms")


if __name__ == "__main__":
    main()
