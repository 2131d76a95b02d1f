"""Where the per-step time outside the UNet graph goes: 25 back-to-back replays of the captured B=2 UNet call vs the
full AnimationSampler step loop (config 2 shape).  python tools/loop_overhead.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.pose_net import PoseNetB200  # noqa: E402
from stableanimator_b200.sampler import AnimationSampler  # noqa: E402
from stableanimator_b200.synth import synth_inputs, synth_posenet_state_dict, synth_unet_state_dict  # noqa: E402
from stableanimator_b200.unet import UNetB200  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    cfg = UNetConfig.full()
    unet = UNetB200(cfg, synth_unet_state_dict(cfg), "cuda")
    pose = PoseNetB200(synth_posenet_state_dict(320), 320, "cuda")
    inp = {k: v.cuda() for k, v in synth_inputs(cfg, 16, 64, 64).items()}
    s = AnimationSampler(unet, pose)
    call = lambda: s(inp["latents"], inp["image_latents"], inp["encoder_hidden_states"], inp["added_time_ids"], inp["pose"])  # noqa: E731
    full = timed(call)
    plan = next(iter(s._plans.values()))
    run = next(iter(plan["runners"].values()))

    def replays():
        for _ in range(25):
            run()
    only = timed(replays)
    one = timed(lambda: s(inp["latents"], inp["image_latents"], inp["encoder_hidden_states"], inp["added_time_ids"],
                          inp["pose"], steps_to_run=1))
    two = timed(lambda: s(inp["latents"], inp["image_latents"], inp["encoder_hidden_states"], inp["added_time_ids"],
                          inp["pose"], steps_to_run=2))
    print(f"1 step: {one:.2f} ms; 2 steps: {two:.2f} ms -> per-video setup (PoseNet, context, noise tiling) "
          f"{2 * one - two:.2f} ms, one step {two - one:.2f} ms; one bare replay: {timed(run, 5):.2f} ms")
    def replays_with_fill():   # one tiny eager kernel in front of every replay, as the step loop has
        for _ in range(25):
            run.t.fill_(1.0)
            run()
    fill = timed(replays_with_fill)
    print(f"25 x (one eager fill_ + graph replay): {fill:.1f} ms ({(fill - only) / 25 * 1e3:.0f} us per step over bare replays)")
    print(f"sampler call (25 steps): {full:.1f} ms; 25 bare graph replays: {only:.1f} ms; difference {full - only:.1f} ms "
          f"({(full - only) / 25 * 1e3:.0f} us per step)")


if __name__ == "__main__":
    main()
