"""CPU/fp32 ORACLE of StableAnimator's temporal VAE decoder -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional restatement (plain torch fp32 ops over a ``state_dict``) of ``TemporalDecoder.forward``
(/root/reference/animation/modules/vae.py:437-479) with its mid / up blocks (:557-658), the video residual
block (:784-949), ``AlphaBlender`` (:951-965), ``Upsample2D`` (:743-781), the mid-block ``Attention``
(:31-111), and of ``decode_latents`` (animation/pipelines/inference_pipeline_animation.py:319-357).
Pinned against the unmodified reference by oracle/make_golden.py (fixture tests/golden/vae_decoder_tiny.pt,
re-checked by tests/test_oracle_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's reference
legs may import this file.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from stableanimator_b200.arch import VaeDecoderConfig, vae_decoder_blocks, vae_encoder_blocks

SD = Dict[str, torch.Tensor]


def _gn(sd: SD, name: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.group_norm(x, 32, sd[f"{name}.weight"], sd[f"{name}.bias"], eps)


def resnet_block_2d(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """ResnetBlock2D.forward, vae.py:815-832 (eps 1e-6, no time embedding)."""
    h = F.conv2d(F.silu(_gn(sd, f"{name}.norm1", x, 1e-6)), sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(sd, f"{name}.norm2", h, 1e-6)), sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"], padding=1)
    if f"{name}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{name}.conv_shortcut.weight"], sd[f"{name}.conv_shortcut.bias"])
    return x + h


def temporal_resnet_block(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """TemporalResnetBlock.forward, vae.py:877-898 on [B, C, F, h, w] (eps 1e-5, Conv3d (3,1,1))."""
    h = F.conv3d(F.silu(_gn(sd, f"{name}.norm1", x, 1e-5)), sd[f"{name}.conv1.weight"], sd[f"{name}.conv1.bias"],
                 padding=(1, 0, 0))
    h = F.conv3d(F.silu(_gn(sd, f"{name}.norm2", h, 1e-5)), sd[f"{name}.conv2.weight"], sd[f"{name}.conv2.bias"],
                 padding=(1, 0, 0))
    return x + h


def spatio_temporal_res_block(sd: SD, name: str, x: torch.Tensor, frames: int) -> torch.Tensor:
    """SpatioTemporalResBlock.forward, vae.py:925-948, with AlphaBlender.forward (:957-965):
    alpha = 1 - sigmoid(mix_factor); out = alpha * spatial + (1 - alpha) * temporal."""
    xs = resnet_block_2d(sd, f"{name}.spatial_res_block", x)
    bf, c, h, w = xs.shape
    x5 = xs.reshape(bf // frames, frames, c, h, w).permute(0, 2, 1, 3, 4)
    xt = temporal_resnet_block(sd, f"{name}.temporal_res_block", x5)
    alpha = 1.0 - torch.sigmoid(sd[f"{name}.time_mixer.mix_factor"])
    out = alpha * x5 + (1.0 - alpha) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def attention(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """Attention.forward, vae.py:60-111: GroupNorm(eps 1e-6) -> q/k/v Linear (bias) -> single-head SDPA
    (scale head_dim^-1/2) -> to_out -> + residual, tokens = h*w per frame."""
    b, c, h, w = x.shape
    t = _gn(sd, f"{name}.group_norm", x.view(b, c, h * w), 1e-6).transpose(1, 2)
    q = F.linear(t, sd[f"{name}.to_q.weight"], sd[f"{name}.to_q.bias"])
    k = F.linear(t, sd[f"{name}.to_k.weight"], sd[f"{name}.to_k.bias"])
    v = F.linear(t, sd[f"{name}.to_v.weight"], sd[f"{name}.to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, sd[f"{name}.to_out.0.weight"], sd[f"{name}.to_out.0.bias"])
    return o.transpose(1, 2).reshape(b, c, h, w) + x


def temporal_decoder(sd: SD, cfg: VaeDecoderConfig, z: torch.Tensor, num_frames: int) -> torch.Tensor:
    """TemporalDecoder.forward, vae.py:437-479: z [B*F, 4, h, w] -> [B*F, 3, 8h, 8w]."""
    x = F.conv2d(z, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    for bname, res, up in vae_decoder_blocks(cfg):
        for j, (rname, _, _) in enumerate(res):
            if bname == "mid_block" and j >= 1:
                x = attention(sd, "mid_block.attentions.0", x)   # MidBlockTemporalDecoder.forward, vae.py:597-611
            x = spatio_temporal_res_block(sd, rname, x, num_frames)
        if up:   # Upsample2D.forward, vae.py:754-781: nearest x2 then conv3x3
            x = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), sd[f"{up}.weight"], sd[f"{up}.bias"], padding=1)
    x = F.conv2d(F.silu(_gn(sd, "conv_norm_out", x, 1e-6)), sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    bf, c, h, w = x.shape
    x5 = x.reshape(bf // num_frames, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    x5 = F.conv3d(x5, sd["time_conv_out.weight"], sd["time_conv_out.bias"], padding=(1, 0, 0))
    return x5.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def decode_latents(sd: SD, cfg: VaeDecoderConfig, latents: torch.Tensor, num_frames: int,
                   decode_chunk_size: int = 4) -> torch.Tensor:
    """InferenceAnimationPipeline.decode_latents (inference_pipeline_animation.py:319-357):
    latents [B, F, 4, h, w] -> frames fp32 [B, 3, F, 8h, 8w], decoded decode_chunk_size frames at a time."""
    z = latents.flatten(0, 1) / cfg.scaling_factor
    frames = []
    for i in range(0, z.shape[0], decode_chunk_size):
        chunk = z[i:i + decode_chunk_size]
        frames.append(temporal_decoder(sd, cfg, chunk, chunk.shape[0]))
    frames = torch.cat(frames, dim=0)
    return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()


def vae_encode_moments(sd: SD, cfg: VaeDecoderConfig, image: torch.Tensor) -> torch.Tensor:
    """AutoencoderKLTemporalDecoder.encode up to the posterior parameters (vae.py:253-256): Encoder.forward
    (vae.py:537-555: conv_in, DownEncoderBlock2D x4 with Downsample2D = F.pad(0,1,0,1) + stride-2 conv (:731-741),
    UNetMidBlock2D (:380-384), GroupNorm + SiLU + conv_out) followed by quant_conv.  image [B, 3, H, W] ->
    moments [B, 8, H/8, W/8] (mean | logvar)."""
    x = F.conv2d(image, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for res, down in vae_encoder_blocks(cfg):
        for name, _, _ in res:
            x = resnet_block_2d(sd, f"encoder.{name}", x)
        if down:
            x = F.conv2d(F.pad(x, (0, 1, 0, 1)), sd[f"encoder.{down}.weight"], sd[f"encoder.{down}.bias"], stride=2)
    x = resnet_block_2d(sd, "encoder.mid_block.resnets.0", x)
    x = attention(sd, "encoder.mid_block.attentions.0", x)
    x = resnet_block_2d(sd, "encoder.mid_block.resnets.1", x)
    x = F.conv2d(F.silu(_gn(sd, "encoder.conv_norm_out", x, 1e-6)), sd["encoder.conv_out.weight"],
                 sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(x, sd["quant_conv.weight"], sd["quant_conv.bias"])


def vae_encode_mode(sd: SD, cfg: VaeDecoderConfig, image: torch.Tensor) -> torch.Tensor:
    """``vae.encode(image).latent_dist.mode()`` (inference_pipeline_animation.py:257; vae.py:183,203-204):
    the mean half of the moments, [B, 4, H/8, W/8]."""
    return vae_encode_moments(sd, cfg, image).chunk(2, dim=1)[0]
