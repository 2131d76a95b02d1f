"""CPU/fp32 ORACLE of StableAnimator's face-token encoder -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional restatement (plain torch fp32 ops over a ``state_dict``) of ``FusionFaceId.forward``
(/root/reference/animation/modules/id_encoder.py:133-148) and the ``FacePerceiver`` /
``PerceiverAttention`` / ``FeedForward`` it calls (:15-100).  Pinned against the unmodified reference by
oracle/make_golden.py (fixture tests/golden/face_fusion.pt, re-checked by tests/test_oracle_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this file.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _ln(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{name}.weight"], sd[f"{name}.bias"], 1e-5)


def _heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    """reshape_tensor, id_encoder.py:8-13: [b, l, h*d] -> [b, h, l, d]."""
    b, l, _ = x.shape
    return x.view(b, l, heads, -1).transpose(1, 2)


def perceiver_attention(sd: SD, name: str, x: torch.Tensor, latents: torch.Tensor, heads: int = 16) -> torch.Tensor:
    """PerceiverAttention.forward, id_encoder.py:33-57 (q/k each scaled by d^-1/4, softmax in fp32)."""
    x = _ln(sd, f"{name}.norm1", x)
    latents = _ln(sd, f"{name}.norm2", latents)
    b, l, _ = latents.shape
    q = F.linear(latents, sd[f"{name}.to_q.weight"])
    k, v = F.linear(torch.cat((x, latents), dim=-2), sd[f"{name}.to_kv.weight"]).chunk(2, dim=-1)
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    scale = 1.0 / math.sqrt(math.sqrt(q.shape[-1]))
    w = torch.softmax(((q * scale) @ (k * scale).transpose(-2, -1)).float(), dim=-1)
    out = (w @ v).permute(0, 2, 1, 3).reshape(b, l, -1)
    return F.linear(out, sd[f"{name}.to_out.weight"])


def feed_forward(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward, id_encoder.py:59-66: LayerNorm -> Linear -> GELU(erf) -> Linear, no biases."""
    h = F.linear(_ln(sd, f"{name}.0", x), sd[f"{name}.1.weight"])
    return F.linear(F.gelu(h), sd[f"{name}.3.weight"])


def fusion_face_id(sd: SD, id_embeds: torch.Tensor, clip_embeds: torch.Tensor, num_tokens: int = 4,
                   depth: int = 4) -> torch.Tensor:
    """FusionFaceId.forward (id_encoder.py:133-148) -> face tokens [B, num_tokens, 1024]."""
    x = F.linear(id_embeds, sd["proj.0.weight"], sd["proj.0.bias"])
    x = F.linear(F.gelu(x), sd["proj.2.weight"], sd["proj.2.bias"])
    latents = _ln(sd, "norm", x.reshape(-1, num_tokens, sd["norm.weight"].shape[0]))
    # FacePerceiver.forward(latents, x = clip_embeds), id_encoder.py:94-100
    xc = F.linear(clip_embeds, sd["fusion_model.proj_in.weight"], sd["fusion_model.proj_in.bias"])
    for i in range(depth):
        latents = perceiver_attention(sd, f"fusion_model.layers.{i}.0", xc, latents) + latents
        latents = feed_forward(sd, f"fusion_model.layers.{i}.1", latents) + latents
    latents = F.linear(latents, sd["fusion_model.proj_out.weight"], sd["fusion_model.proj_out.bias"])
    return _ln(sd, "fusion_model.norm_out", latents)
