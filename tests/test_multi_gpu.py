"""Multi-GPU paths on real devices (skipped on a single-GPU box): the N-rank tiled sampler (units partitioned, one NCCL
all-reduce per Euler step) must give the latents of the single-rank sampler; a model built on cuda:1 must run while
cuda:0 is the current device (the launchers make the tensors' device current)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")


@needs2
def test_two_rank_sampler_equals_single_rank():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(ROOT / "tools" / "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "MULTI_GPU_CHECK_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-1500:])


@needs2
def test_model_on_cuda1_while_cuda0_is_current():
    from oracle import unet_oracle as O
    from stableanimator_b200.arch import UNetConfig
    from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict
    from stableanimator_b200.unet import UNetB200
    from util import rel_l2
    torch.cuda.set_device(0)
    cfg = UNetConfig.tiny()
    sd = synth_unet_state_dict(cfg)
    inp = synth_inputs(cfg, 4, 16, 16)
    sample = torch.cat([inp["latents"] * 3.0, inp["image_latents"][1:2]], dim=2)
    t = torch.tensor(0.977053)
    unet = UNetB200(cfg, sd, "cuda:1")
    out = unet(sample.to("cuda:1"), t, inp["encoder_hidden_states"][1:2].to("cuda:1"), inp["added_time_ids"][1:2].to("cuda:1"))[0]
    assert torch.cuda.current_device() == 0 and out.device.index == 1
    ref = O.unet_forward(sd, cfg, sample, t, inp["encoder_hidden_states"][1:2], inp["added_time_ids"][1:2])
    assert rel_l2(out.float().cpu(), ref) <= 1e-2
