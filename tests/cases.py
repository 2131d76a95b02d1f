"""Rebuild the exact inputs of the golden cases (tests/golden/unet_*.pt, written by
oracle/make_golden.py from the unmodified reference) from the deterministic generators."""
from pathlib import Path

import torch

from stableanimator_b200.arch import UNetConfig
from stableanimator_b200.synth import synth_inputs

GOLD = Path(__file__).resolve().parent / "golden"


def load_golden(name: str):
    g = torch.load(GOLD / f"unet_{name}.pt", map_location="cpu", weights_only=False)
    cfg = UNetConfig.full() if name.startswith("full") else UNetConfig.tiny()
    frames, h, w, cond = g["frames"], g["h"], g["w"], g["cond"]
    inp = synth_inputs(cfg, frames, h, w, g["seed"])
    row = 1 if cond else 0
    sample = torch.cat([inp["latents"].expand(1, frames, 4, h, w) * 3.0,
                        inp["image_latents"][row:row + 1]], dim=2)
    return dict(cfg=cfg, gold=g, sample=sample, t=torch.tensor(g["t"]),
                ehs=inp["encoder_hidden_states"][row:row + 1], ati=inp["added_time_ids"][row:row + 1],
                pose=inp["pose"], pose_latents=g["pose_latents"])
