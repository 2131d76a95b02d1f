#!/usr/bin/env python
"""Benchmark of the denoising hot path (BASELINE.json metric: denoised frames/s at 512x512x16f,
25 Euler steps).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (B200 kernels)
  python bench.py --impl reference [...]                       the reference's own CPU path (oracle port)

A "step" is one full 25-step denoise of the workload through ``AnimationSampler``.  N=1: BASELINE
configs[1] (512x512 -> 64x64 latent, 16 frames, one tile, CFG cond/uncond fused in one B=2 batch).
N>1 (torchrun): weak scaling of the long-video sampler -- N frame tiles of 16 frames (overlap 4,
a video of 16 + 12 (N-1) frames), one tile (both CFG branches) per rank, one NCCL all-reduce of the
weighted noise accumulator per Euler step.  ``value`` = tile frames denoised by all ranks (16 per rank:
the unit of work a rank processes; overlap frames are denoised by both neighbouring tiles and blended,
inference_pipeline_animation.py:656-689) / max-over-ranks time; the unique video frames per second are
reported next to it as ``video_frames_per_s``.
``--config 2|3|4`` selects the other BASELINE.json configs (fixed videos, strong scaling: 576x1024x16f with the CFG
branches on two GPUs; 512x512x64f and 576x1024x128f with the (frame tile, CFG branch) units spread over the ranks);
there ``value`` = the video's frames / max-over-ranks time.
Prints ONE JSON line (rank 0).  Synthetic weights/inputs (stableanimator_b200/synth.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "denoised frames/sec at 512x512x16f 25 steps"
UNIT = "frames/s"
# reference-graph FLOPs of one B=1 UNet call at F=16, 64x64 / 128x72 latent
FLOPS_PER_CALL_64x64x16 = 20.311e12
FLOPS_PER_CALL_128x72x16 = 51.200e12
# BASELINE.json configs[i] -> (latent h, latent w, video frames | None = weak scaling, description)
CONFIGS = {
    1: (64, 64, None, "512x512, 16-frame tiles, CFG cond/uncond fused (B=2)"),
    2: (128, 72, 16, "576x1024 (w x h), 16 frames, CFG branches split over the ranks"),
    3: (64, 64, 64, "512x512, 64-frame long video, tile 16 / overlap 4, (tile, CFG branch) units over the ranks"),
    4: (128, 72, 128, "576x1024 (w x h), 128-frame long video, tile 16 / overlap 4, (tile, CFG branch) units over the ranks"),
}
# average DRAM bytes (read + write) per sa_gemm_tap launch in one B=2 UNet call at 64x64x16, from the ncu pass
# `--metrics dram__bytes_read.sum,dram__bytes_write.sum` over all 334 launches of the family
# (profiles/r1_v8_launches_unet_b2_64x64x16.csv, summary next to it)
GEMM_DRAM_BYTES_PER_LAUNCH = 132.2e6


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------- reference / CPU leg
def cpu_reference_sample(frames: int, hw: int, cfg_name: str = "full", w: int = 0):
    """Time ONE B=1 UNet call of the oracle port (fp32, all host cores) at the workload's shape and
    extrapolate to the 50 calls (25 steps x 2 CFG branches) of a 16-frame denoise."""
    from oracle import unet_oracle as O
    from stableanimator_b200.arch import UNetConfig
    from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = UNetConfig.full() if cfg_name == "full" else UNetConfig.tiny()
    sd = synth_unet_state_dict(cfg)
    w = w or hw
    inp = synth_inputs(cfg, frames, hw, w)
    sample = torch.cat([inp["latents"].expand(1, frames, 4, hw, w), inp["image_latents"][:1]], dim=2)
    t0 = time.perf_counter()
    O.unet_forward(sd, cfg, sample, torch.tensor(1.0), inp["encoder_hidden_states"][:1], inp["added_time_ids"][:1])
    dt = time.perf_counter() - t0
    return dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 16
    cores = os.cpu_count() or 1
    # bounded sample: full 16-frame call when the host has enough cores, else 4 frames (scaled x4)
    sample_frames = 16 if cores >= 32 else 4
    times = []
    for i in range(args.warmup + args.steps):
        dt, cores = cpu_reference_sample(sample_frames, 64, args.size)
        if i >= args.warmup:
            times.append(dt)
        if sum(times) > 150:
            break
    per_call = sum(times) / len(times) * (frames / sample_frames)
    sec_per_video = per_call * 50  # 25 steps x (uncond + cond), inference_pipeline_animation.py:663-685
    v = frames / sec_per_video
    sample = (f"1 B=1 fp32 UNet call of the oracle port, F={sample_frames}, 64x64 latent, scaled x{frames // sample_frames} "
              f"to F=16 and x50 calls (25 steps x 2 CFG branches)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": sec_per_video * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "512x512x16f, 25 Euler steps, CFG (BASELINE configs[1])"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- our arm
def gemm_family_time(unet, ctx, ati, B, F, h, w, reps=5):
    """Device time of the dominant kernel family (sa_gemm_tap, tcgen05) inside one UNet call, measured
    live with CUDA events on the launching stream: replay a CUDA graph of the full call and a graph of
    the same call with the GEMM launches elided (all other kernels still run, on whatever the output
    buffers contain -- none of them has data-dependent control flow); the difference is the in-situ
    GEMM time, with the real cache state and launch overlap of the step.
    Returns (algorithmic flops of the family, family ms, launches, full-call ms)."""
    from stableanimator_b200 import ops
    x_in = torch.zeros((B * F, h, w, unet.cfg.in_channels), device=unet.dev, dtype=torch.bfloat16)
    t = torch.tensor(1.0, device=unet.dev)
    orig = ops.gemm_tap
    flops = []

    def elided(a0, w_, **kw):
        rows = a0.numel() // a0.shape[-1] // (kw.get("in_stride", 1) ** 2)   # output positions
        n = kw.get("n") or w_.shape[0]
        flops.append(2.0 * rows * w_.shape[1] * n)
        if kw.get("out") is not None:
            return kw["out"]
        n_out = n // 2 if kw.get("geglu") else n
        return torch.empty((rows, n_out), device=a0.device,
                           dtype=torch.float32 if kw.get("out_f32") else torch.bfloat16)

    def graph_ms(fn):
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

    run = lambda: unet.forward_nhwc(x_in, t, ctx, ati, B, F, h, w)  # noqa: E731
    full = graph_ms(run)
    ops.gemm_tap = elided
    try:
        rest = graph_ms(run)
    finally:
        ops.gemm_tap = orig
    n_launch = len(flops) // 2  # the elided function ran twice (warm-up + capture)
    return sum(flops) / 2, full - rest, n_launch, full


def run_ours(args):
    import torch.distributed as dist
    from stableanimator_b200 import _lib
    from stableanimator_b200.arch import UNetConfig
    from stableanimator_b200.pose_net import PoseNetB200
    from stableanimator_b200.sampler import AnimationSampler
    from stableanimator_b200.synth import synth_inputs, synth_posenet_state_dict, synth_unet_state_dict
    from stableanimator_b200.unet import UNetB200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference")
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line (rank 0): everything else that lands on fd 1 during the run (NCCL's
    # version banner, library chatter) is sent to stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    _lib.lib()  # fail loudly if the CUDA library is missing

    cfg = UNetConfig.tiny() if args.size == "tiny" else UNetConfig.full()
    lh, lw, fixed_frames, cfg_desc = CONFIGS[args.config]
    if args.latent != 64:
        lh = lw = args.latent
    hw = lh
    tile = 16
    # configs[1]: weak scaling, N tiles of 16 frames (overlap 4); configs[2..4]: one fixed video, strong scaling
    frames = fixed_frames if fixed_frames else tile + (tile - 4) * (world - 1)
    strong = fixed_frames is not None
    unet = UNetB200(cfg, synth_unet_state_dict(cfg), dev)
    pose_net = PoseNetB200(synth_posenet_state_dict(cfg.block_out_channels[0]), cfg.block_out_channels[0], dev)
    inp = synth_inputs(cfg, frames, lh, lw)
    host = {k: inp[k].pin_memory() for k in ("image_latents", "encoder_hidden_states", "added_time_ids", "pose")}
    host["latents"] = inp["latents"][:, :tile].contiguous().pin_memory()
    devt = {k: v.to(dev) for k, v in host.items()}
    sampler = AnimationSampler(unet, pose_net, num_inference_steps=args.euler_steps, tile_size=tile, tile_overlap=4,
                               use_cuda_graph=not args.no_graph)

    def call(src):
        return sampler(src["latents"], src["image_latents"], src["encoder_hidden_states"], src["added_time_ids"],
                       src["pose"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        call(devt)
    barrier()

    # ---- device-resident timing (value)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        call(devt)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _lib.launch_count - n0
    clk = clocks.stop() if rank == 0 else None

    # ---- end-to-end timing through the public API with HOST buffers (H2D + D2H inside)
    out_host = torch.empty((1, frames, 4, lh, lw), dtype=torch.float32).pin_memory()
    call(host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host.copy_(call(host), non_blocking=False)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        tt = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = tt.tolist()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    if rank == 0:
        sustained, burst, hbm, src = peaks()
        # weak: frames denoised per bench step by all ranks (N=1: the 16-frame video); strong: the video's frames
        tile_frames = frames if strong else tile * world
        value = tile_frames / (ms / 1e3)
        e2e = tile_frames / (e2e_ms / 1e3)
        # roofline of the dominant kernel family (tcgen05 tapped GEMM): live CUDA-event pass
        ctx = unet.prepare_context(devt["encoder_hidden_states"])
        g_flops, g_ms, g_n, call_ms = gemm_family_time(unet, ctx, devt["added_time_ids"], 2, tile, lh, lw)
        achieved = g_flops / (g_ms * 1e-3) / 1e12
        from stableanimator_b200 import scheduler as sch
        n_tiles = len(sch.tile_indices(frames, tile, 4))
        calls_per_step = args.euler_steps * 2 * n_tiles  # B=1-equivalent UNet calls per bench step
        step_tflops = None
        per_call = {(64, 64): FLOPS_PER_CALL_64x64x16, (128, 72): FLOPS_PER_CALL_128x72x16}.get((lh, lw))
        if args.size == "full" and per_call:
            step_tflops = calls_per_step * per_call / (ms * 1e-3) / 1e12 / world  # per GPU
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (f"{lw * 8}x{lh * 8}x{frames}f, {args.euler_steps} Euler steps, {cfg_desc}, "
                                    f"tile 16 / overlap 4 (BASELINE configs[{args.config}])"
                                    + ("" if world == 1 else f", {n_tiles} frame tiles over {world} GPUs")),
                       "frames_counted": (f"the video's {frames} frames (fixed video, {2 * n_tiles} UNet units per step)" if strong else
                                          f"tile frames denoised per step by all ranks ({tile} x {world}); "
                                          f"unique video frames = {frames}"),
                       "unet": args.size, "l2": "working set (3.1 GB weights + activations per UNet call) >> 126 MB L2; no flush",
                       "cuda_graph": not args.no_graph},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "video_frames_per_s": frames / (ms / 1e3),
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                         "frac": achieved / sustained,
                         "traffic": GEMM_DRAM_BYTES_PER_LAUNCH if (args.size == "full" and (lh, lw) == (64, 64)) else None,
                         "traffic_unit": "bytes of DRAM read+write per launch (ncu, family average)",
                         "peak_source": src,
                         "kernel": "gemm_tap_kernel (tcgen05 tapped implicit GEMM, all instantiations)",
                         "launches_per_unet_call": g_n, "family_ms_per_unet_call": g_ms,
                         "unet_call_ms": call_ms, "family_share_of_call": g_ms / call_ms,
                         "how": "CUDA-graph replay of one B=2 UNet call minus the same graph with the GEMM launches elided",
                         "whole_step_tflops_reference_graph_per_gpu": step_tflops,
                         "whole_step_frac": None if step_tflops is None else step_tflops / sustained},
        }
        if world == 1 and not args.no_cpu:
            sample_frames = 16 if (os.cpu_count() or 1) >= 32 else 4
            dt, cores = cpu_reference_sample(sample_frames, lh, args.size, lw)
            sec = dt * (16 / sample_frames) * 50
            line["cpu_baseline"] = {"value": 16 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"1 B=1 fp32 oracle UNet call, F={sample_frames}, {lh}x{lw} latent, "
                                              f"scaled to F=16 and x50 calls"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="full", choices=["full", "tiny"], help="tiny = debug-size UNet (not a bench result)")
    ap.add_argument("--latent", type=int, default=64, help="latent height = width (64 = 512x512)")
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS),
                    help="BASELINE.json configs[i]: 1 = the metric's config (default; weak scaling for N > 1), "
                         "2 = 576x1024x16f CFG split (2 GPUs), 3 = 512x512x64f (4 / 8 GPUs), 4 = 576x1024x128f (8 GPUs)")
    ap.add_argument("--euler-steps", type=int, default=25)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
