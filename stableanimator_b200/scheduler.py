"""Host-side scalars of the Euler/EDM sampler: Karras sigma table, continuous timesteps, tile index
lists and blend weights.  Mirrors what ``diffusers.EulerDiscreteScheduler`` (third-party; SVD-xt
config: karras sigmas 0.002..700, v_prediction, continuous timesteps) and
inference_pipeline_animation.py:581,604-616,656-657 compute on the host; the per-element arithmetic
(scale_model_input, CFG, step) runs in the fused sampler kernels."""
from __future__ import annotations

import math
from typing import List


def karras_sigmas(n: int, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0) -> List[float]:
    lo, hi = sigma_min ** (1.0 / rho), sigma_max ** (1.0 / rho)
    out = []
    for i in range(n):
        r = i / (n - 1) if n > 1 else 0.0
        out.append((hi + r * (lo - hi)) ** rho)
    return out + [0.0]


def timesteps(sigmas: List[float]) -> List[float]:
    return [0.25 * math.log(s) for s in sigmas[:-1]]


def init_noise_sigma(sigmas: List[float]) -> float:
    return math.sqrt(max(sigmas) ** 2 + 1.0)


def tile_indices(num_frames: int, tile_size: int, tile_overlap: int) -> List[List[int]]:
    if num_frames < tile_size:
        raise ValueError(f"num_frames ({num_frames}) must be >= tile_size ({tile_size})")
    idx = [[0, *range(i + 1, min(i + tile_size, num_frames))]
           for i in range(0, num_frames - tile_size + 1, tile_size - tile_overlap)]
    if idx[-1][-1] < num_frames - 1:
        idx.append([0, *range(num_frames - tile_size + 1, num_frames)])
    return idx


def tile_weights(tile_size: int) -> List[float]:
    return [min(w, 2.0 - w) for w in ((k + 0.5) * 2.0 / tile_size for k in range(tile_size))]


def partition_units(n_tiles: int, world: int) -> List[List[tuple]]:
    """Static split of the 2*n_tiles (tile, cfg-branch) units over ``world`` ranks: contiguous,
    balanced chunks so a tile's uncond+cond pair stays on one rank whenever possible (B=2 fusion)."""
    units = [(t, b) for t in range(n_tiles) for b in (0, 1)]
    out, start = [], 0
    for r in range(world):
        cnt = len(units) // world + (1 if r < len(units) % world else 0)
        out.append(units[start:start + cnt])
        start += cnt
    return out
