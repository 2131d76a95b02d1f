"""CPU/fp32 ORACLE of the whole ``InferenceAnimationPipeline.__call__`` -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Chains the per-stage oracles in the order of /root/reference/animation/pipelines/inference_pipeline_animation.py:494-706:
``_encode_image`` (clip_oracle, :505) -> face tokens and the [CLIP | 4 face tokens] / zeros layout (face_oracle, :509-540)
-> noise augmentation + ``vae.encode(...).latent_dist.mode()`` (vae_oracle, :542-565) -> ``_get_add_time_ids`` (:567-577)
-> the tiled CFG / Euler loop (sampler_oracle + unet_oracle, :580-699) -> ``decode_latents`` (vae_oracle, :701-702) ->
``tensor2vid`` with output_type "pt" (:136-165, :704).  Every stage is pinned on its own against the unmodified
reference or ``transformers`` (oracle/make_golden.py); the pipeline module itself cannot be imported offline (it needs
diffusers' pipeline classes), so the chain as a whole is "parity unpinned" beyond its stages and the line-by-line
citations below.  Random draws (``randn_tensor`` at :543 and :394) are arguments.
Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this file.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import clip_oracle, face_oracle, sampler_oracle, unet_oracle, vae_oracle


def encode_conditioning(sd: Dict[str, dict], cfgs: Tuple, image01: torch.Tensor, id_embedding: torch.Tensor) -> torch.Tensor:
    """:505-540 -> encoder_hidden_states [2, 1 + 4, C], unconditional (zeros) row first."""
    _, _, ccfg = cfgs
    emb = clip_oracle.encode_image(sd["clip"], ccfg, image01)                                    # :505
    face = face_oracle.fusion_face_id(sd["face"], id_embedding.reshape(1, -1).float(), emb[1:])   # :509-524
    uncond = torch.cat([emb[:1], torch.zeros_like(face)], dim=1)                                 # :526-533
    cond = torch.cat([emb[1:], face], dim=1)                                                     # :535-538
    return torch.cat([uncond, cond])                                                             # :540


def encode_image_latents(sd, cfgs, image01: torch.Tensor, image_noise: torch.Tensor, noise_aug_strength: float,
                         num_frames: int) -> torch.Tensor:
    """:542-565 for an image already at (height, width): [2, F, 4, h, w], unconditional zeros first."""
    _, vcfg, _ = cfgs
    img = (2.0 * image01.float() - 1.0) + noise_aug_strength * image_noise                        # :542-545
    il = vae_oracle.vae_encode_mode(sd["venc"], vcfg, img)                                       # :552-557
    return torch.cat([torch.zeros_like(il), il]).unsqueeze(1).repeat(1, num_frames, 1, 1, 1)      # :262-268, :565


def animation_pipeline(sd: Dict[str, dict], cfgs: Tuple, image01: torch.Tensor, pose: torch.Tensor,
                       id_embedding: torch.Tensor, latents: torch.Tensor, image_noise: torch.Tensor, *,
                       num_frames: int, tile_size: int = 16, tile_overlap: int = 4, num_inference_steps: int = 25,
                       min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0, fps: int = 7,
                       motion_bucket_id: int = 127, noise_aug_strength: float = 0.02, decode_chunk_size: int = 4):
    """Returns (encoder_hidden_states, image_latents, final latents [1, F, 4, h, w], frames [F, 3, H, W] in [0, 1]).
    ``sd``: state dicts {"unet", "pose", "face", "clip", "venc", "vdec"}; ``cfgs`` = (UNetConfig, VaeDecoderConfig,
    ClipVisionConfig); ``pose`` [F, 3, H, W] in [-1, 1] (:619-625)."""
    ucfg, vcfg, _ = cfgs
    ehs = encode_conditioning(sd, cfgs, image01, id_embedding)
    image_latents = encode_image_latents(sd, cfgs, image01, image_noise, noise_aug_strength, num_frames)
    ati = torch.tensor([[fps - 1, motion_bucket_id, noise_aug_strength]] * 2, dtype=torch.float32)  # :507, :567-577
    lat = sampler_oracle.denoise_loop(
        lambda s, t, e, a, pl: unet_oracle.unet_forward(sd["unet"], ucfg, s, t, e, a, pl),
        lambda p: unet_oracle.pose_net(sd["pose"], p), latents, image_latents, ehs, ati, pose,
        num_steps=num_inference_steps, tile_size=tile_size, tile_overlap=tile_overlap,
        min_guidance=min_guidance_scale, max_guidance=max_guidance_scale)                         # :580-699
    frames = vae_oracle.decode_latents(sd["vdec"], vcfg, lat, num_frames, decode_chunk_size)      # :702
    video = (frames[0].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)                                # :704, postprocess "pt"
    return ehs, image_latents, lat, video
