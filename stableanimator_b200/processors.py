"""Attention-processor slots of the reference, backed by the B200 kernels (the operator-level
drop-in boundary).

Same classes, constructor arguments, parameter names and call signature as the reference's
``AnimationAttnProcessor`` (animation/modules/unet.py:1855-1914),
``AnimationIDAttnNormalizedProcessor`` (:1916-2029, owns ``id_to_k`` / ``id_to_v``) and
``XFormersAttnProcessor`` (:1812-1853), so ``unet.set_attn_processor({...})`` of the stock reference
UNet accepts them unchanged (wiring recipe: inference_basic.py:348-377,445)::

    processor(attn, hidden_states[B', S, C], encoder_hidden_states=None) -> Tensor[B', S, C]

``attn`` is the reference ``Attention`` module (unet.py:711-824): the processors read
``attn.to_q/to_k/to_v`` (Linear, no bias), ``attn.to_out[0]`` (Linear + bias) and ``attn.heads``; the
input is not modified, a fresh tensor in the input dtype is returned and the caller adds the residual.
Errors are Python exceptions; there is no fallback to PyTorch attention.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops


def _act_dtype(x: torch.Tensor) -> torch.dtype:
    """fp16 hidden states (the reference's default inference dtype, inference_basic.py:486) run the kernels in fp16;
    anything else (bf16, fp32) runs them in bf16."""
    return torch.float16 if x.dtype == torch.float16 else torch.bfloat16


def _pack(attn, extra: Optional[nn.Module], dev, act=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """16-bit / fp32 device copies of the projection weights, cached on the attention module and
    invalidated when any parameter is replaced or modified in place."""
    params = [attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight, attn.to_out[0].bias]
    if extra is not None:
        params += [extra.id_to_k.weight, extra.id_to_v.weight]
    key = tuple((p.data_ptr(), p._version) for p in params) + (str(dev), str(act))
    cache = getattr(attn, "_sa_b200_pack", None)
    if cache is not None and cache["key"] == key:
        return cache
    bf = lambda t: t.detach().to(dev, act).contiguous()  # noqa: E731
    bsl = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()  # noqa: E731  (sa_small_linear: bf16 weights)
    c = {"key": key, "q": bf(attn.to_q.weight), "k": bf(attn.to_k.weight), "v": bf(attn.to_v.weight),
         "o": bf(attn.to_out[0].weight), "ob": attn.to_out[0].bias.detach().to(dev, torch.float32).contiguous()}
    if attn.to_k.weight.shape[1] == attn.to_q.weight.shape[1]:  # self-attention: fused q|k|v projection
        c["qkv"] = torch.cat([c["q"], c["k"], c["v"]], 0).contiguous()
    if extra is not None:
        c["ik"], c["iv"] = bsl(extra.id_to_k.weight), bsl(extra.id_to_v.weight)
        c["v_sl"], c["o_sl"] = bsl(attn.to_v.weight), bsl(attn.to_out[0].weight)
        c["rowsum"] = c["o"].float().sum(dim=1).contiguous()
    attn.__dict__["_sa_b200_pack"] = c  # plain attribute: not a registered parameter/buffer
    return c


def _tokens(x: torch.Tensor, act=torch.bfloat16) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("stableanimator_b200 processors need CUDA tensors (no CPU fallback)")
    return x.to(act).contiguous().reshape(-1, x.shape[-1])


class AnimationAttnProcessor(nn.Module):
    """Spatial self-attention per frame (unet.py:1860-1914): fused q|k|v GEMM -> tcgen05 flash
    attention (head_dim 64) -> to_out GEMM with bias."""

    def __init__(self, hidden_size=320):
        super().__init__()
        self.hidden_size = hidden_size

    def __call__(self, attn, hidden_states, encoder_hidden_states=None):
        assert hidden_states is not None
        if encoder_hidden_states is not None and encoder_hidden_states is not hidden_states:
            raise NotImplementedError("AnimationAttnProcessor is only used for self-attention (unet.py:535-538)")
        bsz, s, c = hidden_states.shape
        act = _act_dtype(hidden_states)
        p = _pack(attn, None, hidden_states.device, act)
        qkv = ops.gemm_tap(_tokens(hidden_states, act), p["qkv"], f16_from_col=2 * c)  # v block in fp16
        o = ops.attn_spatial(qkv, bsz, s, attn.heads)
        out = ops.gemm_tap(o, p["o"], bias=p["ob"])
        return out.reshape(bsz, s, c).to(hidden_states.dtype)


class AnimationIDAttnNormalizedProcessor(nn.Module):
    """ID-adapter cross-attention with distribution alignment (unet.py:1938-2029)."""

    def __init__(self, hidden_size=320, cross_attention_dim=1024, scale=1.0, num_tokens=4):
        super().__init__()
        self.hidden_size = hidden_size
        assert cross_attention_dim == 1024
        self.scale = scale
        self.id_to_k = nn.Linear(cross_attention_dim, hidden_size, bias=False)
        self.id_to_v = nn.Linear(cross_attention_dim, hidden_size, bias=False)
        self.num_tokens = num_tokens
        assert self.scale == 1.0

    def __call__(self, attn, hidden_states, encoder_hidden_states=None):
        assert hidden_states is not None
        assert encoder_hidden_states is not None
        bsz, s, c = hidden_states.shape
        dev = hidden_states.device
        ehs = encoder_hidden_states.to(dev, torch.float32)
        end = ehs.shape[1] - self.num_tokens
        if end != 1 or self.num_tokens != 4:
            raise NotImplementedError("expected 1 CLIP token followed by 4 face tokens (unet.py:1968-1972)")
        act = _act_dtype(hidden_states)
        p = _pack(attn, self, dev, act)
        q = ops.gemm_tap(_tokens(hidden_states, act), p["q"])
        clip = ehs[:, 0].contiguous()                               # [B', 1024]
        face = ehs[:, 1:].reshape(bsz * 4, -1).contiguous()         # [B'*4, 1024]
        vclip = ops.small_linear(clip, p["v_sl"])                   # softmax over ONE key == V (:1974-1997)
        ik = ops.small_linear(face, p["ik"]).reshape(bsz, 4, c)     # :1999-2000
        iv = ops.small_linear(face, p["iv"]).reshape(bsz, 4, c)
        wv = ops.small_linear(vclip, p["o_sl"])
        stats = torch.empty((bsz, 2), device=dev, dtype=torch.float64)
        ip = ops.attn_id(q, ik, iv, bsz, 1, s, attn.heads, stats)   # :2010 (+ Sum, Sum^2 for :2019)
        rs = torch.empty((bsz,), device=dev, dtype=torch.float32)
        rb = torch.empty((bsz, c), device=dev, dtype=torch.float32)
        ops.attn_id_finalize(stats, vclip, wv, p["rowsum"], p["ob"], bsz, 1, s, c, rs, rb)  # :2018-2020
        out = ops.gemm_tap(ip, p["o"], row_scale=rs, row_bias=rb, rows_per_group=s)      # :2022-2024
        return out.reshape(bsz, s, c).to(hidden_states.dtype)


class XFormersAttnProcessor:
    """Temporal attention processor (unet.py:1820-1853): ``hidden_states`` is ``[B*S, F, C]``.
    Self-attention over the F <= 16 frames, or cross-attention against ONE context token
    (``encoder_hidden_states [B*S, 1, 1024]``, the only shape the reference produces, unet.py:640-654)."""

    def __init__(self, attention_op=None):
        self.attention_op = attention_op
        assert self.attention_op is None

    def __call__(self, attn, hidden_states, encoder_hidden_states=None):
        bsz, f, c = hidden_states.shape
        act = _act_dtype(hidden_states)
        p = _pack(attn, None, hidden_states.device, act)
        if encoder_hidden_states is None:
            qkv = ops.gemm_tap(_tokens(hidden_states, act), p["qkv"])
            o = ops.attn_temporal(qkv, bsz, f, 1, attn.heads)
            out = ops.gemm_tap(o, p["o"], bias=p["ob"])
            return out.reshape(bsz, f, c).to(hidden_states.dtype)
        if encoder_hidden_states.shape[1] != 1:
            raise NotImplementedError("temporal cross-attention is only defined for a single context token")
        ctx = _tokens(encoder_hidden_states, act)                    # [B*S, 1024]
        v = ops.gemm_tap(ctx, p["v"])                                # softmax over one key == V
        o = ops.gemm_tap(v, p["o"], bias=p["ob"])                    # [B*S, C]
        return o[:, None, :].expand(bsz, f, c).contiguous().to(hidden_states.dtype)
