"""CPU/fp32 ORACLE of the CLIP image-conditioning path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional restatement (plain torch fp32 ops over a ``state_dict``) of ``InferenceAnimationPipeline._encode_image``
(/root/reference/animation/pipelines/inference_pipeline_animation.py:192-244):

  * ``_resize_with_antialiasing`` (:99-125) with its helpers ``_gaussian`` / ``_gaussian_blur2d`` / ``_filter2d`` /
    ``_compute_padding`` (:22-96): separable Gaussian blur (sigma = max((factor - 1) / 2, 0.001), odd kernel of
    ~2*2 sigma taps, reflect padding) followed by bicubic ``interpolate(align_corners=True)``;
  * the ``CLIPImageProcessor`` normalisation (:213-220) -- third-party ``transformers`` (reference pins 4.35.2,
    requirements.txt; this image has 5.5.0), restated as ``(x - mean) / std`` per channel;
  * ``CLIPVisionModelWithProjection(image).image_embeds`` (:224) -- third-party ``transformers``: patch conv (no bias) +
    class token + learned positions, ``pre_layrnorm``, N pre-norm encoder layers (q/k/v/out with bias, softmax(q k^T
    d^-1/2) v, fc1 - exact GELU - fc2), ``post_layernorm`` of the class token, bias-free ``visual_projection``.

Pinned by oracle/make_golden.py against the reference's own ``_resize_with_antialiasing`` and against
``transformers.CLIPVisionModelWithProjection`` / ``CLIPImageProcessor`` executed here on the same synthetic weights
(fixture tests/golden/clip_vision_tiny.pt, re-checked on CPU by tests/test_oracle_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this file.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def blur_geometry(in_hw: Tuple[int, int], out_hw: Tuple[int, int]):
    """sigma and (odd) kernel size per axis, inference_pipeline_animation.py:100-120."""
    sig, ks = [], []
    for n_in, n_out in zip(in_hw, out_hw):
        s = max((n_in / n_out - 1.0) / 2.0, 0.001)
        k = int(max(2.0 * 2 * s, 3))
        sig.append(s)
        ks.append(k + 1 if k % 2 == 0 else k)
    return tuple(sig), tuple(ks)


def gaussian_taps(k: int, sigma: float) -> torch.Tensor:
    """Normalised Gaussian window (:67-80) for an odd ``k``."""
    x = torch.arange(k, dtype=torch.float32) - k // 2
    g = torch.exp(-x.pow(2.0) / (2 * torch.tensor(sigma, dtype=torch.float32).pow(2.0)))
    return g / g.sum()


def _blur_axis(x: torch.Tensor, taps: torch.Tensor, dim: int) -> torch.Tensor:
    """Depthwise 1-D correlation along ``dim`` (2 = rows, 3 = columns) with reflect padding (:44-64)."""
    k = taps.numel()
    pad = (k - 1) // 2
    b, c, h, w = x.shape
    if dim == 3:
        xp = F.pad(x, (pad, k - 1 - pad, 0, 0), mode="reflect")
        wgt = taps.view(1, 1, 1, k).expand(c, 1, 1, k)
    else:
        xp = F.pad(x, (0, 0, pad, k - 1 - pad), mode="reflect")
        wgt = taps.view(1, 1, k, 1).expand(c, 1, k, 1)
    return F.conv2d(xp, wgt, groups=c)


def resize_with_antialiasing(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """``_resize_with_antialiasing(input, size)`` (:99-125): blur along x, then y, then bicubic align_corners."""
    (sy, sx), (ky, kx) = blur_geometry(x.shape[-2:], size)
    x = _blur_axis(x, gaussian_taps(kx, sx), 3)
    x = _blur_axis(x, gaussian_taps(ky, sy), 2)
    return F.interpolate(x, size=size, mode="bicubic", align_corners=True)


def clip_pixel_values(image01: torch.Tensor, cfg) -> torch.Tensor:
    """``_encode_image`` up to the encoder call (:205-220): image in [0, 1] -> normalised [B, 3, S, S]."""
    x = image01.float() * 2.0 - 1.0
    x = resize_with_antialiasing(x, (cfg.image_size, cfg.image_size))
    x = (x + 1.0) / 2.0
    mean = torch.tensor(cfg.image_mean, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(cfg.image_std, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def _ln(sd: SD, name: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{name}.weight"], sd[f"{name}.bias"], eps)


def clip_image_embeds(sd: SD, cfg, pixel_values: torch.Tensor) -> torch.Tensor:
    """CLIPVisionModelWithProjection.forward(pixel_values).image_embeds -> [B, projection_dim]."""
    v = "vision_model"
    b = pixel_values.shape[0]
    c, heads = cfg.hidden_size, cfg.num_attention_heads
    d = c // heads
    patches = F.conv2d(pixel_values, sd[f"{v}.embeddings.patch_embedding.weight"], stride=cfg.patch_size)
    x = patches.flatten(2).transpose(1, 2)                                   # [B, np, C]
    cls = sd[f"{v}.embeddings.class_embedding"].expand(b, 1, c)
    x = torch.cat([cls, x], dim=1) + sd[f"{v}.embeddings.position_embedding.weight"][None]
    x = _ln(sd, f"{v}.pre_layrnorm", x, cfg.layer_norm_eps)
    n = x.shape[1]
    for l in range(cfg.num_hidden_layers):
        p = f"{v}.encoder.layers.{l}"
        h = _ln(sd, f"{p}.layer_norm1", x, cfg.layer_norm_eps)
        q, k, val = (F.linear(h, sd[f"{p}.self_attn.{m}.weight"], sd[f"{p}.self_attn.{m}.bias"])
                     .view(b, n, heads, d).transpose(1, 2) for m in ("q_proj", "k_proj", "v_proj"))
        w = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, dim=-1)
        o = (w @ val).transpose(1, 2).reshape(b, n, c)
        x = x + F.linear(o, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
        h = _ln(sd, f"{p}.layer_norm2", x, cfg.layer_norm_eps)
        h = F.gelu(F.linear(h, sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"]))
        x = x + F.linear(h, sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"])
    pooled = _ln(sd, f"{v}.post_layernorm", x[:, 0], cfg.layer_norm_eps)
    return F.linear(pooled, sd["visual_projection.weight"])


def encode_image(sd: SD, cfg, image01: torch.Tensor, do_classifier_free_guidance: bool = True) -> torch.Tensor:
    """``_encode_image`` (:192-244) for one video per prompt: [B, 3, H, W] in [0, 1] -> [2B | B, 1, proj] with the
    zero (unconditional) embeddings first (:232-241)."""
    emb = clip_image_embeds(sd, cfg, clip_pixel_values(image01, cfg)).unsqueeze(1)
    return torch.cat([torch.zeros_like(emb), emb]) if do_classifier_free_guidance else emb
