"""Deterministic synthetic weights and inputs (there are no checkpoints or datasets offline).

Every tensor is generated from ``crc32(key) ^ seed`` on the CPU, so the build container, the GPU
box, the oracle and the product all see bit-identical values without shipping any file.  Values
are rounded to bf16-representable fp32 ("fp32 oracle on bf16-rounded weights").
Magnitudes follow PyTorch's default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for Linear/Conv,
norm weights near 1) so activations stay in a realistic range through the 38 residual stages.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch

from .arch import (ClipVisionConfig, UNetConfig, VaeDecoderConfig, clip_vision_param_shapes, face_param_shapes,
                   posenet_param_shapes, unet_param_shapes, vae_decoder_param_shapes, vae_encoder_param_shapes)


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "mix_factor":
        t = torch.full(shape, 0.5) + 0.25 * (torch.rand(shape, generator=g) - 0.5)
    elif leaf == "scale":  # PoseNet output scale (pose_net.py:38)
        t = torch.full(shape, 2.0)
    elif "norm" in key.rsplit(".", 2)[-2] or (len(shape) == 1 and key.rsplit(".", 2)[-2] == "0" and ".layers." in key):
        # LayerNorm / GroupNorm affine parameters (incl. the LayerNorm at index 0 of the perceiver FeedForward)
        if leaf == "weight":
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        else:
            t = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
    elif leaf == "weight":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = fan_in ** -0.5
        t = (torch.rand(shape, generator=g) * 2 - 1) * bound
    else:  # bias of a Linear / Conv: same bound as its weight; fan_in unknown here -> small
        t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return bf16_round(t)


def synth_unet_state_dict(cfg: UNetConfig, seed: int = 23123134) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, s, seed) for k, s in unet_param_shapes(cfg).items()}


def synth_posenet_state_dict(noise_latent_channels: int = 320, seed: int = 23123134):
    return {k: synth_tensor("pose_net." + k, s, seed)
            for k, s in posenet_param_shapes(noise_latent_channels).items()}


def synth_vae_decoder_state_dict(cfg: VaeDecoderConfig, seed: int = 23123134):
    """TemporalDecoder weights (animation/modules/vae.py:387-435), keys as in its ``state_dict``."""
    return {k: synth_tensor("vae.decoder." + k, s, seed) for k, s in vae_decoder_param_shapes(cfg).items()}


def synth_vae_encoder_state_dict(cfg: VaeDecoderConfig, seed: int = 23123134):
    """``encoder.*`` + ``quant_conv.*`` of AutoencoderKLTemporalDecoder (vae.py:236-247,482-556)."""
    return {k: synth_tensor("vae." + k, s, seed) for k, s in vae_encoder_param_shapes(cfg).items()}


def synth_clip_vision_state_dict(cfg: ClipVisionConfig, seed: int = 23123134):
    """CLIPVisionModelWithProjection weights, keys as in its ``state_dict``."""
    return {k: synth_tensor("image_encoder." + k, s, seed) for k, s in clip_vision_param_shapes(cfg).items()}


def synth_image(h: int, w: int, seed: int = 23123134, batch: int = 1) -> torch.Tensor:
    """Reference image in [-1, 1] (smooth low-frequency pattern + noise), [B, 3, H, W]."""
    g = _gen("input.image", seed)
    base = torch.nn.functional.interpolate(torch.randn(batch, 3, max(h // 16, 1), max(w // 16, 1), generator=g),
                                           size=(h, w), mode="bilinear", align_corners=False)
    return bf16_round((0.6 * base + 0.1 * torch.randn(batch, 3, h, w, generator=g)).clamp(-1, 1))


def synth_vae_latents(frames: int, h: int, w: int, seed: int = 23123134) -> torch.Tensor:
    """Denoised latents [1, F, 4, h, w] in the sampler's scale (|z| ~ 1; the decoder divides by 0.18215)."""
    return bf16_round(torch.randn(1, frames, 4, h, w, generator=_gen("input.vae_latents", seed)))


def synth_face_state_dict(seed: int = 23123134, **dims):
    """FusionFaceId weights (animation/modules/id_encoder.py:102-131), keys as in its ``state_dict``."""
    return {k: synth_tensor("face_encoder." + k, s, seed) for k, s in face_param_shapes(**dims).items()}


def synth_face_inputs(batch: int = 1, seed: int = 23123134, id_dim: int = 512, clip_dim: int = 1024):
    """ArcFace-like identity embedding [B, 512] and CLIP image embedding [B, 1, 1024] (ranges of
    id_encoder.py:135-136)."""
    ide = torch.randn(batch, id_dim, generator=_gen("input.id_embeds", seed))
    clip = torch.randn(batch, 1, clip_dim, generator=_gen("input.clip_embeds", seed)) * 1.5
    return bf16_round(ide), bf16_round(clip)


def synth_inputs(cfg: UNetConfig, frames: int, h: int, w: int, seed: int = 23123134,
                 pose_channels: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic conditioning for one video.  ``h, w`` are latent sizes.

    Returns fp32 CPU tensors, bf16-rounded: ``latents`` [1,F,4,h,w] (unit variance; the sampler
    scales by init_noise_sigma), ``image_latents`` [2,F,4,h,w] (row 0 = uncond zeros),
    ``encoder_hidden_states`` [2,5,1024] (row 0 zeros; row 1 = CLIP token + 4 face tokens),
    ``added_time_ids`` [2,3], ``pose`` [F,3,8h,8w] in [-1,1].
    """
    def n(name, *shape):
        return torch.randn(*shape, generator=_gen("input." + name, seed))

    out = {}
    out["latents"] = bf16_round(n("latents", 1, frames, 4, h, w))
    img = n("image_latents", 1, 1, 4, h, w).expand(1, frames, 4, h, w) * 8.0
    out["image_latents"] = bf16_round(torch.cat([torch.zeros_like(img), img], dim=0).contiguous())
    clip = n("clip", 1, 1, cfg.cross_attention_dim) * 1.5
    face = n("face", 1, cfg.num_tokens, cfg.cross_attention_dim) * 3.0
    ehs = torch.cat([clip, face], dim=1)
    out["encoder_hidden_states"] = bf16_round(torch.cat([torch.zeros_like(ehs), ehs], dim=0))
    out["added_time_ids"] = torch.tensor([[6.0, 127.0, 0.02]] * 2)
    # pose maps: -1 background with a few bright line segments (reference mean is -0.994)
    H, W = 8 * h, 8 * w
    pose = torch.full((frames, 3, H, W), -1.0)
    g = _gen("input.pose", seed)
    for f in range(frames):
        for _ in range(6):
            y = int(torch.randint(0, H, (1,), generator=g))
            x0 = int(torch.randint(0, max(W - 8, 1), (1,), generator=g))
            ln = int(torch.randint(4, max(W // 4, 5), (1,), generator=g))
            col = torch.rand(3, generator=g) * 2 - 1
            pose[f, :, y:y + 2, x0:x0 + ln] = col[:, None, None]
    out["pose"] = bf16_round(pose)
    return out
