"""CPU/fp32 ORACLE of the denoising loop around the UNet -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ``InferenceAnimationPipeline.__call__`` steps 4-8
(/root/reference/animation/pipelines/inference_pipeline_animation.py:580-699) and the third-party
``diffusers.EulerDiscreteScheduler`` it drives (not vendored in the reference; version unpinned,
requirements.txt:1; SVD-xt scheduler_config: karras sigmas, sigma 0.002..700, v_prediction,
continuous timesteps).  The scheduler restatement is pinned by the constants the reference prints:
``init_noise_sigma == 700.000732`` (:405) and the 25 timesteps ``1.637770 .. -1.553652`` (:633-639),
checked in tests/test_scheduler_kat.py.  The ``step`` formula itself has no in-tree pin
("parity unpinned" for that formula; it follows diffusers' published v-prediction Euler step).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import torch

# timesteps printed by the reference run, inference_pipeline_animation.py:633-639
REFERENCE_TIMESTEPS_25 = [
    1.637770, 1.575531, 1.510996, 1.443990, 1.374316, 1.301752, 1.226049, 1.146922, 1.064048,
    0.977053, 0.885506, 0.788904, 0.686657, 0.578063, 0.462282, 0.338294, 0.204848, 0.060379,
    -0.097098, -0.270160, -0.462234, -0.678018, -0.924202, -1.210778, -1.553652]
REFERENCE_INIT_NOISE_SIGMA = 700.000732  # :405


def karras_sigmas(n: int, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0) -> torch.Tensor:
    """Karras et al. (2022) eq. 5 as used by EulerDiscreteScheduler._convert_to_karras; appends 0."""
    ramp = torch.linspace(0, 1, n, dtype=torch.float64)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    s = (hi + ramp * (lo - hi)) ** rho
    return torch.cat([s, torch.zeros(1, dtype=torch.float64)]).to(torch.float32)


def timesteps_from_sigmas(sigmas: torch.Tensor) -> torch.Tensor:
    """timestep_type == 'continuous' with v_prediction: t = 0.25 * ln(sigma)."""
    return 0.25 * torch.log(sigmas[:-1])


def init_noise_sigma(sigmas: torch.Tensor) -> float:
    """timestep_spacing 'leading': sqrt(sigma_max^2 + 1)."""
    return float((sigmas.max() ** 2 + 1) ** 0.5)


def scale_model_input(x: torch.Tensor, sigma: float) -> torch.Tensor:
    return x / ((sigma ** 2 + 1) ** 0.5)


def euler_step_vpred(model_out: torch.Tensor, sample: torch.Tensor, sigma: float, sigma_next: float) -> torch.Tensor:
    """EulerDiscreteScheduler.step, prediction_type 'v_prediction', s_churn = 0 (no noise)."""
    sample = sample.float()
    x0 = model_out.float() * (-sigma / (sigma ** 2 + 1) ** 0.5) + sample / (sigma ** 2 + 1)
    d = (sample - x0) / sigma
    return sample + d * (sigma_next - sigma)


def tile_indices(num_frames: int, tile_size: int, tile_overlap: int) -> List[List[int]]:
    """inference_pipeline_animation.py:613-616: every tile is [0] + tile_size-1 consecutive frames."""
    idx = [[0, *range(i + 1, min(i + tile_size, num_frames))]
           for i in range(0, num_frames - tile_size + 1, tile_size - tile_overlap)]
    if idx[-1][-1] < num_frames - 1:
        idx.append([0, *range(num_frames - tile_size + 1, num_frames)])
    return idx


def tile_weights(tile_size: int) -> torch.Tensor:
    """:656-657 triangular blend weights."""
    w = (torch.arange(tile_size, dtype=torch.float32) + 0.5) * 2.0 / tile_size
    return torch.minimum(w, 2 - w)


@torch.no_grad()
def denoise_loop(unet_fn: Callable, pose_fn: Optional[Callable], latents: torch.Tensor,
                 image_latents: torch.Tensor, encoder_hidden_states: torch.Tensor,
                 added_time_ids: torch.Tensor, pose: Optional[torch.Tensor], *, num_steps: int = 25,
                 tile_size: int = 16, tile_overlap: int = 4, min_guidance: float = 3.0,
                 max_guidance: float = 3.0, steps_to_run: Optional[int] = None) -> torch.Tensor:
    """inference_pipeline_animation.py:580-699 in fp32.

    ``unet_fn(sample[1,T,8,h,w], t, ehs[1,5,1024], ati[1,3], pose_latents|None) -> [1,T,4,h,w]``;
    ``pose_fn(pose[T,3,H,W]) -> [T,320,h,w]``.  ``latents`` [1,tile,4,h,w] is UNIT-variance noise
    (scaled here by init_noise_sigma, then tiled along frames like :597); ``image_latents``
    [2,F,4,h,w]; ``encoder_hidden_states`` [2,5,1024]; ``added_time_ids`` [2,3]; ``pose`` [F,3,H,W].
    Returns the final latents [1,F,4,h,w].
    """
    num_frames = image_latents.shape[1]
    sig = karras_sigmas(num_steps)
    ts = timesteps_from_sigmas(sig)
    lat = latents.float() * init_noise_sigma(sig)
    lat = lat.repeat(1, num_frames // tile_size + 1, 1, 1, 1)[:, :num_frames]                # :597
    dev = image_latents.device   # CPU in the tests; cuda for bench.py's eager-GPU baseline leg
    g = torch.linspace(min_guidance, max_guidance, num_frames).reshape(1, num_frames, 1, 1, 1).to(dev)  # :604-607
    indices = tile_indices(num_frames, tile_size, tile_overlap)
    weight = tile_weights(tile_size).to(dev)
    n = num_steps if steps_to_run is None else steps_to_run
    for i in range(n):
        sigma, sigma_next = float(sig[i]), float(sig[i + 1])
        x = scale_model_input(torch.cat([lat] * 2), sigma)                                     # :644-647
        x = torch.cat([x, image_latents.float()], dim=2)                                       # :650
        noise = torch.zeros_like(image_latents, dtype=torch.float32)
        cnt = torch.zeros(num_frames, device=dev)
        for idx in indices:                                                                    # :659
            pl = pose_fn(pose[idx].float()) if (pose_fn is not None and pose is not None) else None
            eu = unet_fn(x[:1, idx], ts[i], encoder_hidden_states[:1], added_time_ids[:1], None)
            noise[:1, idx] += eu * weight[:, None, None, None]                                 # :673
            ec = unet_fn(x[1:, idx], ts[i], encoder_hidden_states[1:], added_time_ids[1:], pl)
            noise[1:, idx] += ec * weight[:, None, None, None]                                 # :685
            cnt[idx] += weight                                                                 # :687
        noise = noise / cnt[:, None, None, None]                                               # :689
        eps = noise[:1] + g * (noise[1:] - noise[:1])                                          # :694-696
        lat = euler_step_vpred(eps, lat, sigma, sigma_next)                                    # :699
    return lat
