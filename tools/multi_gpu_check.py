"""torchrun --nproc-per-node N tools/multi_gpu_check.py
Every rank runs the tiled sampler twice on the tiny UNet: once as a member of the N-rank group (units
partitioned, NCCL all-reduce per step) and once alone (world 1); the final latents must agree."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.pose_net import PoseNetB200  # noqa: E402
from stableanimator_b200.sampler import AnimationSampler  # noqa: E402
from stableanimator_b200.synth import synth_inputs, synth_posenet_state_dict, synth_unet_state_dict  # noqa: E402
from stableanimator_b200.unet import UNetB200  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = UNetConfig.tiny()
    frames, h, w, tile, overlap = 22, 8, 8, 4, 1
    unet = UNetB200(cfg, synth_unet_state_dict(cfg), f"cuda:{local}")
    pose = PoseNetB200(synth_posenet_state_dict(cfg.block_out_channels[0]), cfg.block_out_channels[0], f"cuda:{local}")
    inp = synth_inputs(cfg, frames, h, w)
    args = (inp["latents"][:, :tile].contiguous(), inp["image_latents"], inp["encoder_hidden_states"],
            inp["added_time_ids"], inp["pose"])
    multi = AnimationSampler(unet, pose, num_inference_steps=6, tile_size=tile, tile_overlap=overlap)
    out_multi = multi(*args).clone()
    single = AnimationSampler(unet, pose, num_inference_steps=6, tile_size=tile, tile_overlap=overlap)
    single.rank, single.world = 0, 1
    out_single = single(*args).clone()
    err = ((out_multi - out_single).norm() / out_single.norm()).item()
    errs = [None] * dist.get_world_size()
    dist.all_gather_object(errs, err)
    if dist.get_rank() == 0:
        plan = multi._plans[next(iter(multi._plans))]
        print(f"world={dist.get_world_size()} tiles={len(plan['tiles'])} rel-L2(multi vs single) per rank: "
              + ", ".join(f"{e:.2e}" for e in errs))
        assert max(errs) < 1e-5, errs
        print("MULTI_GPU_CHECK_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
