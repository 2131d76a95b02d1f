"""Time the full-size temporal VAE decoder: decode_latents of 16 frames at 64x64 latent (512x512 frames)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200.arch import VaeDecoderConfig  # noqa: E402
from stableanimator_b200.synth import synth_vae_decoder_state_dict, synth_vae_latents  # noqa: E402
from stableanimator_b200.vae_decoder import TemporalDecoderB200  # noqa: E402

cfg = VaeDecoderConfig.full()
frames, h, w = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64
dec = TemporalDecoderB200(cfg, synth_vae_decoder_state_dict(cfg), "cuda")
lat = synth_vae_latents(frames, h, w).cuda()
out = dec.decode_latents(lat, frames, 4)
torch.cuda.synchronize()
t0 = time.perf_counter()
out = dec.decode_latents(lat, frames, 4)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"decode_latents {frames} frames {8 * h}x{8 * w}: {dt * 1e3:.1f} ms ({frames / dt:.1f} frames/s), "
      f"out {tuple(out.shape)} finite={bool(torch.isfinite(out).all())} peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
