"""``InferenceAnimationPipeline.__call__`` (animation/pipelines/inference_pipeline_animation.py:443-710) end to end
on the B200 modules: the caller of the hot path.  Same argument names, defaults and step order as the reference:

  1. ``_encode_image`` (:505)                      -> :class:`ClipImageEncoderB200`
  2. face tokens + conditioning layout (:509-540)  -> :class:`FusionFaceIdB200`, ``[CLIP | 4 face tokens]``, uncond = zeros
  3. image noise augmentation + VAE encode (:542-565) -> :class:`VaeEncoderB200` (``latent_dist.mode()``)
  4. ``_get_add_time_ids`` (:567-577), noise latents (:584-597), guidance ramp (:604-611)
  5. the denoising loop (:613-699)                 -> :class:`AnimationSampler` (the hot path)
  6. ``decode_latents`` + ``tensor2vid`` (:701-706)   -> :class:`TemporalDecoderB200`

Host work that is not arithmetic on tensors of the path stays in Python / PIL exactly as in the reference (PIL resize
of the input image by ``VaeImageProcessor.preprocess``, uint8 -> float conversion of the pose frames, output
formatting).  Random numbers come from ``torch.randn`` with the caller's ``generator`` (the reference uses diffusers'
``randn_tensor``); pass ``latents`` / ``image_noise`` explicitly for bit-reproducible runs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .sampler import AnimationSampler


def _to_float_image(image) -> torch.Tensor:
    """PIL image(s) / uint8 arrays -> fp32 [B, 3, H, W] in [0, 1] (pil_to_numpy + numpy_to_pt, :205-206)."""
    if isinstance(image, torch.Tensor):
        return image.float() if image.dim() == 4 else image.float()[None]
    items = image if isinstance(image, (list, tuple)) else [image]
    arr = np.stack([np.asarray(im.convert("RGB") if hasattr(im, "convert") else im) for im in items])
    return torch.from_numpy(arr.astype(np.float32) / 255.0).permute(0, 3, 1, 2).contiguous()


class AnimationPipelineB200:
    def __init__(self, vae_encoder, vae_decoder, image_encoder, unet, pose_net, face_encoder):
        self.vae_encoder, self.vae_decoder, self.image_encoder = vae_encoder, vae_decoder, image_encoder
        self.unet, self.pose_net, self.face_encoder = unet, pose_net, face_encoder
        self.vae_scale_factor = 8
        self.num_tokens = 4
        self.dev = unet.dev
        self._samplers: Dict[tuple, AnimationSampler] = {}

    # ------------------------------------------------------------------ pieces of __call__
    def check_inputs(self, image, height: int, width: int) -> None:
        """:359-371."""
        if not isinstance(image, torch.Tensor) and not hasattr(image, "convert") and not isinstance(image, (list, tuple)):
            raise ValueError(f"`image` has to be a torch.Tensor, a PIL image or a list of PIL images, got {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    def _preprocess(self, image, height: int, width: int) -> torch.Tensor:
        """``VaeImageProcessor.preprocess`` (:542): resize to (height, width) and map [0, 1] -> [-1, 1]."""
        if not isinstance(image, torch.Tensor):
            items = image if isinstance(image, (list, tuple)) else [image]
            from PIL import Image
            image = [im.resize((width, height), resample=Image.LANCZOS) for im in items]
        x = _to_float_image(image)
        if tuple(x.shape[-2:]) != (height, width):
            raise ValueError(f"tensor images must already be {height}x{width}, got {tuple(x.shape[-2:])}")
        return 2.0 * x - 1.0

    def _pose_frames(self, image_pose) -> torch.Tensor:
        """:619-625: frames of pose drawings -> fp32 [F, 3, H, W] in [-1, 1]."""
        if isinstance(image_pose, torch.Tensor):
            return image_pose.float()
        frames = [torch.from_numpy(np.array(p)).float() / 127.5 - 1 for p in image_pose]
        return torch.stack(frames, 0).permute(0, 3, 1, 2).contiguous()

    def _sampler(self, steps, tile_size, tile_overlap, gmin, gmax) -> AnimationSampler:
        key = (steps, tile_size, tile_overlap, float(gmin), float(gmax))
        if key not in self._samplers:
            self._samplers[key] = AnimationSampler(self.unet, self.pose_net, num_inference_steps=steps,
                                                   tile_size=tile_size, tile_overlap=tile_overlap,
                                                   min_guidance_scale=gmin, max_guidance_scale=gmax)
        return self._samplers[key]

    def encode_conditioning(self, image01: torch.Tensor, id_embedding) -> torch.Tensor:
        """:505-540 -> encoder_hidden_states fp32 [2, 1 + 4, 1024] (row 0 = unconditional zeros)."""
        emb = self.image_encoder.encode_image(image01, do_classifier_free_guidance=True)       # [2, 1, C], zeros first
        ide = torch.as_tensor(np.asarray(id_embedding.cpu()) if isinstance(id_embedding, torch.Tensor) else id_embedding)
        ide = ide.reshape(1, -1).to(self.dev, torch.float32)                                 # :509-510
        face = self.face_encoder(ide, emb[1:])                                               # [1, 4, C]
        uncond = torch.cat([emb[:1], torch.zeros_like(face)], dim=1)
        cond = torch.cat([emb[1:], face], dim=1)
        return torch.cat([uncond, cond])

    def encode_image_latents(self, image, height: int, width: int, num_frames: int, noise_aug_strength: float,
                             image_noise: torch.Tensor) -> torch.Tensor:
        """:542-565 -> image_latents fp32 [2, F, 4, h, w] (row 0 = unconditional zeros)."""
        img = self._preprocess(image, height, width).to(self.dev)
        img = img + noise_aug_strength * image_noise.to(self.dev, torch.float32)                # :545
        il = self.vae_encoder.encode_mode(img).float()                                           # [1, 4, h, w]
        return torch.cat([torch.zeros_like(il), il]).unsqueeze(1).repeat(1, num_frames, 1, 1, 1)

    # ------------------------------------------------------------------ __call__
    @torch.no_grad()
    def __call__(self, image, image_pose, height: int = 576, width: int = 1024, num_frames: Optional[int] = None,
                 tile_size: int = 16, tile_overlap: int = 4, num_inference_steps: int = 25,
                 min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0, fps: int = 7,
                 motion_bucket_id: int = 127, noise_aug_strength: float = 0.02, image_only_indicator: bool = False,
                 decode_chunk_size: Optional[int] = None, num_videos_per_prompt: int = 1,
                 generator: Optional[torch.Generator] = None, latents: Optional[torch.Tensor] = None,
                 reference_image_id_ante_embedding=None, output_type: str = "pil", return_dict: bool = True,
                 image_noise: Optional[torch.Tensor] = None):
        if num_videos_per_prompt != 1:
            raise ValueError("num_videos_per_prompt must be 1 (as in the reference)")
        if max_guidance_scale < 1.0:
            raise ValueError("classifier-free guidance is always on in the reference (max_guidance_scale >= 1)")
        if reference_image_id_ante_embedding is None:
            raise ValueError("reference_image_id_ante_embedding (the 512-d face identity embedding) is required")
        pose = self._pose_frames(image_pose)
        num_frames = int(num_frames) if num_frames is not None else pose.shape[0]
        if pose.shape[0] != num_frames or tuple(pose.shape[-2:]) != (height, width):
            raise ValueError(f"image_pose must hold {num_frames} frames of {height}x{width}")
        decode_chunk_size = decode_chunk_size if decode_chunk_size is not None else num_frames   # :494
        self.check_inputs(image, height, width)
        image01 = _to_float_image(image)
        if image01.shape[0] != 1:
            raise ValueError("one reference image per call (the reference UNet asserts batch size 1)")
        ehs = self.encode_conditioning(image01.to(self.dev), reference_image_id_ante_embedding)
        if image_noise is None:
            image_noise = torch.randn((1, 3, height, width), generator=generator, dtype=torch.float32,
                                      device=generator.device if generator is not None else "cpu")
        image_latents = self.encode_image_latents(image, height, width, num_frames, noise_aug_strength, image_noise)
        ati = torch.tensor([[fps - 1, motion_bucket_id, noise_aug_strength]] * 2, dtype=torch.float32)  # :507, :567
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        if latents is None:                                                                     # :584-595
            latents = torch.randn((1, tile_size, 4, h, w), generator=generator, dtype=torch.float32,
                                  device=generator.device if generator is not None else "cpu")
        elif tuple(latents.shape) != (1, tile_size, 4, h, w):
            raise ValueError(f"latents must be [1, {tile_size}, 4, {h}, {w}]")
        sampler = self._sampler(num_inference_steps, tile_size, tile_overlap, min_guidance_scale, max_guidance_scale)
        lat = sampler(latents, image_latents, ehs, ati, pose)                                   # [1, F, 4, h, w]
        if output_type == "latent":
            return lat
        frames = self.vae_decoder.decode_latents(lat, num_frames, decode_chunk_size)            # [1, 3, F, H, W]
        return self.tensor2vid(frames, output_type)

    @staticmethod
    def tensor2vid(video: torch.Tensor, output_type: str = "pt") -> List:
        """:136-165 with ``VaeImageProcessor.postprocess``: per batch element ``[F, 3, H, W]`` in [0, 1] ("pt"),
        ``[F, H, W, 3]`` numpy ("np") or a list of PIL images ("pil")."""
        if output_type not in ("pt", "np", "pil"):
            raise ValueError(f"unknown output_type {output_type!r}")
        outs = []
        for vid in video:
            x = (vid.permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)
            if output_type != "pt":
                x = x.cpu().permute(0, 2, 3, 1).float().numpy()
                if output_type == "pil":
                    from PIL import Image
                    x = [Image.fromarray((f * 255).round().astype("uint8")) for f in x]
            outs.append(x)
        return outs
