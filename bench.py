#!/usr/bin/env python
"""Benchmark of the denoising hot path (BASELINE.json metric: denoised frames/s at 512x512x16f,
25 Euler steps).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (B200 kernels)
  python bench.py --impl reference [...]                       the reference's own CPU path (oracle port)

Our line also carries: ``roofline`` (dominant kernel family, live CUDA-event measurement, plus ``memory_bound`` GB/s of
the GroupNorm / LayerNorm / CFG+Euler kernels), ``gpu_eager_baseline`` (N = 1: the reference graph in eager bf16 PyTorch
on the same GPU -- the north star's >= 8x denominator), ``cpu_baseline`` (N = 1: one CFG step of BASELINE configs[0] on the
host cores) and ``long_video`` (BASELINE configs[4], 576x1024x128f, strong scaling over the N ranks).

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
# average DRAM bytes (read + write) per GEMM-family launch (sa_gemm_tap + sa_ff_fused) in one B=2 UNet call at 64x64x16,
# re-measured at round-2 HEAD with the ncu pass `--metrics dram__bytes_read.sum,dram__bytes_write.sum` over the 313 tapped-GEMM
# launches of the call (profiles/r2_v10_launches_unet_b2_64x64x16.csv, summary next to it; r2_v3: 102.7 MB; round 1: 132.2 MB over 343)
GEMM_DRAM_BYTES_PER_LAUNCH = 102.5e6


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
# reference-graph FLOPs of one B=1 UNet call at the CPU sample's shape (BASELINE.md section 3, configs[0])
FLOPS_PER_CALL_32x32x8 = 2.364e12


class CpuCfgStep:
    """BASELINE.md section 4 / BASELINE.json configs[0]: ONE full CFG step of the reference's loop on the host cores --
    two B=1 fp32 UNet calls (uncond, cond), PoseNet, the CFG combine and the Euler update
    (inference_pipeline_animation.py:641-699 restated in oracle/sampler_oracle.py) -- at 256x256 (32x32 latent), 8 frames,
    full-width UNet, ``torch.set_num_threads(os.cpu_count())``.  The reference module itself cannot travel to the GPU
    box (Python + absent dependencies), so the arithmetic is the oracle port that is pinned to it (kind = "port")."""

    def __init__(self, size: str = "full", frames: int = 8, hw: int = 32, w: int = 0):
        from oracle import sampler_oracle as SO
        from oracle import unet_oracle as O
        from stableanimator_b200.arch import UNetConfig
        from stableanimator_b200.synth import synth_inputs, synth_posenet_state_dict, synth_unet_state_dict
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        cfg = UNetConfig.full() if size == "full" else UNetConfig.tiny()
        self.frames, self.h, self.w = frames, hw, (w or hw)
        sd = synth_unet_state_dict(cfg)
        psd = synth_posenet_state_dict(cfg.block_out_channels[0])
        inp = synth_inputs(cfg, frames, self.h, self.w)
        unet_fn = lambda x, t, e, a, pl: O.unet_forward(sd, cfg, x, t, e, a, pl)  # noqa: E731
        pose_fn = lambda pm: O.pose_net(psd, pm)  # noqa: E731
        self._step = lambda: SO.denoise_loop(unet_fn, pose_fn, inp["latents"], inp["image_latents"],  # noqa: E731
                                             inp["encoder_hidden_states"], inp["added_time_ids"], inp["pose"],
                                             num_steps=25, tile_size=frames, tile_overlap=min(4, frames - 1), steps_to_run=1)

    def seconds(self) -> float:
        t0 = time.perf_counter()
        self._step()
        return time.perf_counter() - t0

    def to_metric(self, sec_per_cfg_step: float) -> float:
        """frames/s of the metric's workload (512x512x16f, 25 steps) extrapolated from the sample by the reference
        graph's algorithmic FLOPs (20.311 TFLOP per B=1 call at 64x64x16 vs 2.364 at 32x32x8)."""
        if (self.frames, self.h, self.w) == (8, 32, 32):
            scale = FLOPS_PER_CALL_64x64x16 / FLOPS_PER_CALL_32x32x8
        else:   # debug shapes: tokens (the quadratic attention term is ignored)
            scale = (16 * 64 * 64) / (self.frames * self.h * self.w)
        return 16.0 / (sec_per_cfg_step * scale * 25)

    def describe(self) -> str:
        return (f"one CFG step (2 B=1 fp32 UNet calls + PoseNet + CFG + Euler) of the oracle port at {self.w * 8}x{self.h * 8}, "
                f"{self.frames} frames (BASELINE configs[0]); scaled to 512x512x16f x 25 steps by the reference-graph FLOP "
                f"ratio 20.311 / 2.364 per call")


def cpu_reference_sample(frames: int, hw: int, cfg_name: str = "full", w: int = 0):
    """Seconds of one CFG step of the oracle port at an arbitrary (debug) shape, and the threads used."""
    leg = CpuCfgStep(cfg_name, frames, hw, w)
    return leg.seconds(), leg.cores


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port), every bench step = one bounded
    sample (a full CFG step at BASELINE configs[0]); exactly --warmup untimed and --steps timed samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    leg = CpuCfgStep(args.size) if args.size == "full" else CpuCfgStep("tiny", 4, 16)
    for _ in range(args.warmup):
        leg.seconds()
    times = [leg.seconds() for _ in range(max(args.steps, 1))]
    sec = sum(times) / len(times)
    v = leg.to_metric(sec)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "512x512x16f, 25 Euler steps, CFG (BASELINE configs[1])"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": leg.cores, "kind": "port", "sample": leg.describe(),
                             "sec_per_cfg_step": sec, "frames_per_cfg_step_second": leg.frames / sec},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU eager leg
def gpu_eager_baseline(dev, lh, lw, frames, euler_steps, size):
    """The north star's >= 8x denominator: the reference graph in eager PyTorch on the SAME GPU in bf16 -- the
    fp32-pinned oracle restatement with weights and activations in bfloat16 (what ``.to('cuda', torch.bfloat16)`` does to
    the reference modules, inference_basic.py:486-503; cuDNN / cuBLAS / SDPA kernels picked by torch), driven exactly
    like inference_pipeline_animation.py:641-699: per Euler step two B=1 UNet calls (uncond, cond), PoseNet, CFG, Euler.
    The reference modules themselves cannot travel to the GPU box, hence kind = "port".  Timed with CUDA events:
    one warm-up step, then the full loop."""
    from oracle import sampler_oracle as SO
    from oracle import unet_oracle as O
    from stableanimator_b200.arch import UNetConfig
    from stableanimator_b200.synth import synth_inputs, synth_posenet_state_dict, synth_unet_state_dict
    cfg = UNetConfig.full() if size == "full" else UNetConfig.tiny()
    bf = torch.bfloat16
    sd = {k: v.to(dev, bf) for k, v in synth_unet_state_dict(cfg).items()}
    psd = {k: v.to(dev, bf) for k, v in synth_posenet_state_dict(cfg.block_out_channels[0]).items()}
    inp = {k: v.to(dev) for k, v in synth_inputs(cfg, frames, lh, lw).items()}
    unet_fn = lambda x, t, e, a, pl: O.unet_forward(sd, cfg, x.to(bf), t.to(dev), e.to(bf), a.to(bf),  # noqa: E731
                                                    None if pl is None else pl.to(bf)).float()
    pose_fn = lambda pm: O.pose_net(psd, pm.to(bf))  # noqa: E731

    def loop(n):
        return SO.denoise_loop(unet_fn, pose_fn, inp["latents"], inp["image_latents"], inp["encoder_hidden_states"],
                               inp["added_time_ids"], inp["pose"], num_steps=euler_steps, tile_size=16, tile_overlap=4,
                               steps_to_run=n)
    with torch.no_grad():
        loop(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop(euler_steps)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    del sd, psd, inp
    torch.cuda.empty_cache()
    return ms


# ----------------------------------------------------------------------------------------- our arm
def gemm_family_time(unet, ctx, ati, B, F, h, w, reps=5):
    """Device time of the dominant kernel family (sa_gemm_tap, tcgen05) inside one UNet call, measured
    live with CUDA events on the launching stream: replay a CUDA graph of the full call and a graph of
    the same call with the GEMM launches elided (all other kernels still run, on whatever the output
    buffers contain -- none of them has data-dependent control flow); the difference is the in-situ
    GEMM time, with the real cache state and launch overlap of the step.
    Returns (algorithmic flops of the family, family ms, launches, full-call ms, {family: ms})."""
    from stableanimator_b200 import profiling
    x_in = torch.zeros((B * F, h, w, unet.cfg.in_channels), device=unet.dev, dtype=torch.bfloat16)
    t = torch.tensor(1.0, device=unet.dev)
    flops = []
    run = lambda: unet.forward_nhwc(x_in, t, ctx, ati, B, F, h, w)  # noqa: E731
    res = profiling.family_times(run, ("gemm_tap", "groupnorm", "layernorm", "attn_spatial"), reps, flops)
    n_launch = len(flops) // 2  # the elided function ran twice (warm-up + capture)
    return sum(flops) / 2, res["gemm_tap"], n_launch, res["full_ms"], res


def long_video_record(unet, pose_net, dev, world, rank, euler_steps, size):
    """BASELINE configs[4] -- 576x1024 (w x h), 128 frames, tile 16 / overlap 4: 11 frame tiles x 2 CFG branches = 22
    UNet units per Euler step spread over the ranks (strong scaling; the north star's >= 0.8 @ 8 GPUs claim).  One
    warm-up Euler step (captures the per-unit CUDA graphs), then ONE full denoise, CUDA-event timed, max over ranks."""
    import torch.distributed as dist
    from stableanimator_b200 import scheduler as sch
    from stableanimator_b200.arch import UNetConfig
    from stableanimator_b200.sampler import AnimationSampler
    from stableanimator_b200.synth import synth_inputs
    cfg = UNetConfig.full() if size == "full" else UNetConfig.tiny()
    lh, lw, frames = (128, 72, 128) if size == "full" else (16, 8, 40)
    inp = synth_inputs(cfg, frames, lh, lw)
    src = {k: inp[k].to(dev) for k in ("image_latents", "encoder_hidden_states", "added_time_ids", "pose")}
    src["latents"] = inp["latents"][:, :16].contiguous().to(dev)
    sampler = AnimationSampler(unet, pose_net, num_inference_steps=euler_steps, tile_size=16, tile_overlap=4)
    call = lambda n: sampler(src["latents"], src["image_latents"], src["encoder_hidden_states"],  # noqa: E731
                             src["added_time_ids"], src["pose"], steps_to_run=n)
    call(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(euler_steps)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
    # the one collective of the path, timed on its own: NCCL all-reduce (sum) of the fp32 [2, F, 4, h, w] noise accumulator
    ar_ms = None
    if world > 1:
        buf = torch.zeros((2, frames, 4, lh, lw), device=dev, dtype=torch.float32)
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([a0.elapsed_time(a1) / 10], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ar_ms = tt.item()
    n_tiles = len(sch.tile_indices(frames, 16, 4))
    units = 2 * n_tiles
    rounds = -(-units // world)
    rec = {"workload": f"{lw * 8}x{lh * 8}x{frames}f, {euler_steps} Euler steps, tile 16 / overlap 4 (BASELINE configs[4])",
           "scaling": "strong", "n_gpus": world, "ms_per_denoise": ms, "video_frames_per_s": frames / (ms / 1e3),
           "units_per_step": units, "partition_ceiling": units / (rounds * world),
           "timed": "1 full denoise after a 1-step warm-up, CUDA events, max over ranks"}
    if ar_ms is not None:
        nbytes = 2 * frames * 4 * lh * lw * 4
        rec["exchange"] = {"collective": "ncclAllReduce(sum) of the fp32 noise accumulator, once per Euler step",
                           "bytes_per_step": nbytes, "ms_per_step_standalone": ar_ms,
                           "share_of_denoise": euler_steps * ar_ms / ms,
                           "algbw_GBps": nbytes / (ar_ms * 1e-3) / 1e9}
    if size == "full":
        rec["whole_step_tflops_reference_graph_per_gpu"] = (euler_steps * units * FLOPS_PER_CALL_128x72x16 / (ms * 1e-3)
                                                            / 1e12 / world)
    del sampler, src
    torch.cuda.empty_cache()
    return rec


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

    # ---- BASELINE configs[4] next to the metric's line at every N (strong scaling of the long-video sampler)
    long_video = None
    if args.config == 1 and args.latent == 64 and not args.no_long_video:
        long_video = long_video_record(unet, pose_net, dev, world, rank, args.euler_steps, args.size)

    if rank == 0:
        sustained, burst, hbm, src = peaks()
        # weak: frames denoised per bench step by all ranks (N=1: the 16-frame video); strong: the video's frames
        tile_frames = frames if strong else tile * world
        value = tile_frames / (ms / 1e3)
        e2e = tile_frames / (e2e_ms / 1e3)
        # roofline of the dominant kernel family (tcgen05 tapped GEMM): live CUDA-event pass
        ctx = unet.prepare_context(devt["encoder_hidden_states"])
        g_flops, g_ms, g_n, call_ms, fam = gemm_family_time(unet, ctx, devt["added_time_ids"], 2, tile, lh, lw)
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
                         "kernel": "gemm_tap_persist_kernel + ff_fused_kernel (tcgen05 tapped implicit GEMM family, all instantiations)",
                         "launches_per_unet_call": g_n, "family_ms_per_unet_call": g_ms,
                         "unet_call_ms": call_ms, "family_share_of_call": g_ms / call_ms,
                         "how": "CUDA-graph replay of one B=2 UNet call minus the same graph with the GEMM-family launches (sa_gemm_tap, sa_ff_fused) elided",
                         "whole_step_tflops_reference_graph_per_gpu": step_tflops,
                         "whole_step_frac": None if step_tflops is None else step_tflops / sustained},
        }
        line["roofline"]["family_ms_per_unet_call_all"] = {k: v for k, v in fam.items() if k != "full_ms"}
        if long_video is not None:
            line["long_video"] = long_video
        # the optional legs below run on rank 0 only and never take the headline down with them: a failure is reported
        # in the line instead of losing it
        def leg(name, fn):
            try:
                fn()
            except Exception as e:  # noqa: BLE001
                line[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

        def leg_memory_bound():
            # memory-bound kernels as HBM GB/s (north star): standalone launches at the L0 shape of a B=2 call, inputs
            # rotating over > 126 MB; plus the in-situ family times against SURVEY 8d's algorithmic bytes per call
            from stableanimator_b200 import profiling
            mb = profiling.memory_bound_gbs(dev, hbm_gbs=hbm)
            for name, gb in (("groupnorm", 2 * 4.42), ("layernorm", 2 * 5.17)):   # GB per B=2 call, unfused
                mb[f"{name}_in_situ"] = {"ms_per_unet_call": fam[name], "algorithmic_GB_per_call": gb,
                                         "GB/s": gb / fam[name] * 1e3, "frac": gb / fam[name] * 1e3 / hbm,
                                         "note": "SURVEY 8d bytes of the UNFUSED graph / time of the launches that remain: "
                                                 "statistics passes and LayerNorm sites fused into GEMM kernels move no "
                                                 "bytes of their own, so the fraction can exceed 1"}
            line["roofline"]["memory_bound"] = mb

        def leg_eager():
            eager_ms = gpu_eager_baseline(dev, lh, lw, frames, args.euler_steps, args.size)
            line["gpu_eager_baseline"] = {
                "value": frames / (eager_ms / 1e3), "unit": UNIT, "ms_per_denoise": eager_ms, "dtype": "bf16",
                "kind": "port", "what": "oracle/unet_oracle.py + oracle/sampler_oracle.py in eager PyTorch on this GPU, bf16 "
                                        "weights and activations, two B=1 UNet calls + PoseNet per Euler step "
                                        "(inference_pipeline_animation.py:641-699)",
                "speedup": value / (frames / (eager_ms / 1e3)), "speedup_e2e": e2e / (frames / (eager_ms / 1e3)),
                "ms_per_cfg_step_eager": eager_ms / args.euler_steps, "ms_per_cfg_step_ours": ms / args.euler_steps}

        def leg_cpu():
            cl = CpuCfgStep(args.size) if args.size == "full" else CpuCfgStep("tiny", 4, 16)
            cl.seconds()                                     # 1 warm-up
            ts = sorted(cl.seconds() for _ in range(3))      # median of 3
            line["cpu_baseline"] = {"value": cl.to_metric(ts[1]), "unit": UNIT, "cores": cl.cores, "kind": "port",
                                    "sample": cl.describe() + "; median of 3 after 1 warm-up",
                                    "sec_per_cfg_step": ts[1], "frames_per_cfg_step_second": cl.frames / ts[1]}

        if world == 1 and args.size == "full" and (lh, lw) == (64, 64):
            leg("memory_bound", leg_memory_bound)
        if world == 1 and not args.no_eager:
            leg("gpu_eager_baseline", leg_eager)
        if world == 1 and not args.no_cpu:
            leg("cpu_baseline", leg_cpu)
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
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg (N = 1)")
    ap.add_argument("--no-long-video", action="store_true", help="skip the BASELINE configs[4] long_video sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
