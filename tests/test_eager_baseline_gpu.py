"""Eager-PyTorch CUDA baseline of the hot path (north_star: ">= 8x single-GPU speed-up over the reference's own
eager PyTorch CUDA path at 512x512x16f").  The reference itself cannot run on the GPU box (it is not there and its
dependencies are not installable), so the stand-in is the oracle -- the fp32-verified restatement of the reference
graph out of plain torch ops -- executed on cuda in bf16 (autocast: cuBLAS / cuDNN / SDPA kernels picked by torch),
two B=1 calls per step exactly like inference_pipeline_animation.py:663-685.  Reported, and sanity-bounded."""
import json
import time
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


def _time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_speedup_over_eager_torch_bf16():
    from oracle import unet_oracle as O
    from stableanimator_b200.arch import UNetConfig
    from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict
    from stableanimator_b200.unet import UNetB200

    cfg = UNetConfig.full()
    frames, hw = 16, 64
    sd = synth_unet_state_dict(cfg)
    inp = synth_inputs(cfg, frames, hw, hw)
    sample = torch.cat([inp["latents"].expand(1, frames, 4, hw, hw), inp["image_latents"][1:2]], dim=2)
    t = torch.tensor(1.0)
    ehs, ati = inp["encoder_hidden_states"][1:2], inp["added_time_ids"][1:2]

    sd_gpu = {k: v.cuda().bfloat16() for k, v in sd.items()}

    def eager():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return O.unet_forward(sd_gpu, cfg, sample.cuda().bfloat16(), t.cuda(), ehs.cuda().bfloat16(), ati.cuda())

    eager_ms = _time(eager)  # one B=1 call
    del sd_gpu
    torch.cuda.empty_cache()

    unet = UNetB200(cfg, sd, "cuda")
    ctx = unet.prepare_context(inp["encoder_hidden_states"].cuda())
    x = torch.randn(2 * frames, hw, hw, 8, device="cuda").bfloat16()
    atig = inp["added_time_ids"].cuda()
    tg = t.cuda()
    run = lambda: unet.forward_nhwc(x, tg, ctx, atig, 2, frames, hw, hw)  # noqa: E731
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    ours_ms = _time(g.replay, reps=5)  # one fused B=2 call = both CFG branches
    speedup = 2.0 * eager_ms / ours_ms
    rec = {"eager_torch_bf16_ms_per_b1_call": eager_ms, "ours_ms_per_b2_call": ours_ms,
           "speedup_per_cfg_step": speedup, "shape": "F=16, 64x64 latent, full UNet",
           "eager": "oracle/unet_oracle.py on cuda under torch.autocast(bfloat16)"}
    print(json.dumps(rec))
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "eager_baseline.json").write_text(json.dumps(rec))
    assert speedup > 1.0
