"""bench.py on CPU: the reference arm (`--impl reference`, oracle port on the host cores) must print one JSON line with
the contract's keys -- run here on the debug-size UNet so the suite stays fast; and our arm must refuse to run without
a CUDA device instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--size", "tiny", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cpu_sample_supports_non_square_latents():
    sys.path.insert(0, str(ROOT))
    import bench
    dt, cores = bench.cpu_reference_sample(2, 8, "tiny", 16)
    assert dt > 0 and cores >= 1


def test_cpu_leg_scales_config0_by_the_reference_flop_ratio():
    sys.path.insert(0, str(ROOT))
    import bench
    leg = bench.CpuCfgStep.__new__(bench.CpuCfgStep)
    leg.frames, leg.h, leg.w = 8, 32, 32
    # 1 s per CFG step at 256x256x8f -> 25 steps x 20.311 / 2.364 seconds for the 16-frame 512x512 video
    assert abs(leg.to_metric(1.0) - 16.0 / (25 * 20.311 / 2.364)) < 1e-9


def test_our_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
