"""CPU/fp32 ORACLE of StableAnimator's denoising UNet -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (plain torch fp32 ops over a ``state_dict``) of
``UNetSpatioTemporalConditionModel.forward`` and everything it calls in
/root/reference/animation/modules/unet.py, each function citing the lines it follows.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import it; the product path (stableanimator_b200/) never does.

Parity pin: oracle/make_golden.py runs the UNMODIFIED reference (imported through oracle/stubs) on
the same synthetic weights/inputs and asserts this file matches it to fp32 round-off; the resulting
vectors are committed under tests/golden/ and re-checked by tests/test_oracle_golden.py.
Third-party pieces that are not in the reference tree (xformers attention, diffusers' Euler
scheduler) are restated from their published definitions -- see oracle/sampler_oracle.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from stableanimator_b200.arch import ResSpec, TstSpec, UNetConfig, build_stages

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------ small pieces
def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """get_timestep_embedding, unet.py:365-390: [cos | sin] after the half swap, fp32."""
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t[:, None].float() * freq[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def linear(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))


def timestep_mlp(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """TimestepEmbedding.forward, unet.py:429-436: linear_1 -> SiLU -> linear_2."""
    return linear(sd, f"{name}.linear_2", F.silu(linear(sd, f"{name}.linear_1", x)))


def group_norm(sd: SD, name: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.group_norm(x, 32, sd[f"{name}.weight"], sd[f"{name}.bias"], eps)


def layer_norm(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{name}.weight"], sd[f"{name}.bias"], 1e-5)


def alpha_blend(sd: SD, name: str, x_spatial: torch.Tensor, x_temporal: torch.Tensor) -> torch.Tensor:
    """AlphaBlender, unet.py:447-476 with image_only_indicator == 0 (forced at unet.py:288):
    alpha = sigmoid(mix_factor), one scalar for every element."""
    alpha = torch.sigmoid(sd[f"{name}.mix_factor"])
    return alpha * x_spatial + (1.0 - alpha) * x_temporal


def heads_split(x: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention.head_to_batch_dim, unet.py:782-788: [B,S,H*d] -> [B*H,S,d]."""
    b, s, c = x.shape
    return x.reshape(b, s, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, s, c // heads)


def heads_merge(x: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention.batch_to_head_dim, unet.py:776-780."""
    bh, s, d = x.shape
    return x.reshape(bh // heads, heads, s, d).permute(0, 2, 1, 3).reshape(bh // heads, s, d * heads)


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """xformers.ops.memory_efficient_attention semantics (third-party, unpinned; README.md:69):
    softmax(q k^T / sqrt(d)) v -- attn.scale == d^-0.5 (unet.py:737) at every call site."""
    w = torch.softmax(q @ k.transpose(-1, -2) * (q.shape[-1] ** -0.5), dim=-1)
    return w @ v


def feed_forward(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward/GEGLU, unet.py:854-870: proj -> value * gelu_erf(gate) -> Linear."""
    h = linear(sd, f"{name}.net.0.proj", x)
    val, gate = h.chunk(2, dim=-1)
    return linear(sd, f"{name}.net.2", val * F.gelu(gate))


# ------------------------------------------------------------------------------------ processors
def attn_self(sd: SD, name: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """AnimationAttnProcessor.__call__ (unet.py:1860-1914) and the self-attention case of
    XFormersAttnProcessor.__call__ (unet.py:1820-1853): q/k/v from x, heads of 64, to_out."""
    q, k, v = (heads_split(linear(sd, f"{name}.{p}", x), heads) for p in ("to_q", "to_k", "to_v"))
    return linear(sd, f"{name}.to_out.0", heads_merge(sdpa(q, k, v), heads))


def attn_cross(sd: SD, name: str, x: torch.Tensor, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    """XFormersAttnProcessor.__call__ with encoder_hidden_states (temporal attn2, unet.py:1832-1848)."""
    q = heads_split(linear(sd, f"{name}.to_q", x), heads)
    k = heads_split(linear(sd, f"{name}.to_k", ctx), heads)
    v = heads_split(linear(sd, f"{name}.to_v", ctx), heads)
    return linear(sd, f"{name}.to_out.0", heads_merge(sdpa(q, k, v), heads))


def attn_id_normalized(sd: SD, name: str, x: torch.Tensor, ehs: torch.Tensor, heads: int,
                       num_tokens: int = 4) -> torch.Tensor:
    """AnimationIDAttnNormalizedProcessor.__call__, unet.py:1938-2029."""
    q = heads_split(linear(sd, f"{name}.to_q", x), heads)
    end = ehs.shape[1] - num_tokens                                   # :1968
    clip, face = ehs[:, :end], ehs[:, end:]                           # :1969-1972
    k = heads_split(linear(sd, f"{name}.to_k", clip), heads)          # :1974
    v = heads_split(linear(sd, f"{name}.to_v", clip), heads)
    h = heads_merge(sdpa(q, k, v), heads)                             # :1990,1997
    ik = heads_split(linear(sd, f"{name}.processor.id_to_k", face), heads)   # :1999
    iv = heads_split(linear(sd, f"{name}.processor.id_to_v", face), heads)
    ip = heads_merge(sdpa(q, ik, iv), heads)                          # :2010,2017
    mean_l, std_l = h.mean(dim=(1, 2), keepdim=True), h.std(dim=(1, 2), keepdim=True)     # :2018
    mean_i, std_i = ip.mean(dim=(1, 2), keepdim=True), ip.std(dim=(1, 2), keepdim=True)   # :2019
    ip = (ip - mean_i) * (std_l / (std_i + 1e-5)) + mean_l            # :2020
    return linear(sd, f"{name}.to_out.0", h + ip)                     # :2022-2024


# ------------------------------------------------------------------------------------ blocks
def spatial_block(sd: SD, name: str, x: torch.Tensor, ehs: torch.Tensor, heads: int) -> torch.Tensor:
    """BasicTransformerBlock.forward, unet.py:525-556.  x [B*F,S,C], ehs [B*F,5,1024]."""
    x = attn_self(sd, f"{name}.attn1", layer_norm(sd, f"{name}.norm1", x), heads) + x
    x = attn_id_normalized(sd, f"{name}.attn2", layer_norm(sd, f"{name}.norm2", x), ehs, heads) + x
    return feed_forward(sd, f"{name}.ff", layer_norm(sd, f"{name}.norm3", x)) + x


def temporal_block(sd: SD, name: str, x: torch.Tensor, frames: int, time_ctx: torch.Tensor,
                   heads: int) -> torch.Tensor:
    """TemporalBasicTransformerBlock.forward, unet.py:919-958.  x [B*F,S,C] -> same."""
    bf, s, c = x.shape
    b = bf // frames
    x = x.reshape(b, frames, s, c).permute(0, 2, 1, 3).reshape(b * s, frames, c)      # :931-933
    x = feed_forward(sd, f"{name}.ff_in", layer_norm(sd, f"{name}.norm_in", x)) + x   # :935-938
    x = attn_self(sd, f"{name}.attn1", layer_norm(sd, f"{name}.norm1", x), heads) + x  # :940-942
    x = attn_cross(sd, f"{name}.attn2", layer_norm(sd, f"{name}.norm2", x), time_ctx, heads) + x
    x = feed_forward(sd, f"{name}.ff", layer_norm(sd, f"{name}.norm3", x)) + x        # :950-952
    return x.reshape(b, s, frames, c).permute(0, 2, 1, 3).reshape(bf, s, c)          # :954-956


def transformer_spatio_temporal(sd: SD, t: TstSpec, x: torch.Tensor, ehs: torch.Tensor,
                                frames: int, num_tokens: int = 4) -> torch.Tensor:
    """TransformerSpatioTemporalModel.forward, unet.py:623-708.  x [B*F,C,h,w], ehs [B*F,5,1024]."""
    bf, c, h, w = x.shape
    b = bf // frames
    end = ehs.shape[1] - num_tokens
    tc = ehs[:, :end]                                                            # :640-641
    tc_first = tc.reshape(b, frames, -1, tc.shape[-1])[:, 0]                     # :644-646
    time_ctx = tc_first[:, None].broadcast_to(b, h * w, tc.shape[-2], tc.shape[-1])
    time_ctx = time_ctx.reshape(b * h * w, -1, tc.shape[-1])                     # :649-653
    residual = x
    y = group_norm(sd, f"{t.name}.norm", x, 1e-6)                                # :575,657
    y = y.permute(0, 2, 3, 1).reshape(bf, h * w, c)                              # :661
    y = linear(sd, f"{t.name}.proj_in", y)                                       # :664
    pos = torch.arange(frames, device=x.device).repeat(b, 1).reshape(-1)         # :666-668
    emb = timestep_mlp(sd, f"{t.name}.time_pos_embed",
                       timestep_embedding(pos, c).to(x.dtype))[:, None, :]      # :669-677 (cast :674)
    y = spatial_block(sd, f"{t.name}.transformer_blocks.0", y, ehs, t.heads)     # :681-684
    y_mix = temporal_block(sd, f"{t.name}.temporal_transformer_blocks.0", y + emb, frames,
                           time_ctx, t.heads)                                    # :686-692
    y = alpha_blend(sd, f"{t.name}.time_mixer", y, y_mix)                        # :693-697
    y = linear(sd, f"{t.name}.proj_out", y)                                      # :700
    y = y.reshape(bf, h, w, c).permute(0, 3, 1, 2)                               # :702
    return y + residual                                                          # :705


def resnet_block_2d(sd: SD, name: str, x: torch.Tensor, temb: torch.Tensor, eps: float) -> torch.Tensor:
    """ResnetBlock2D.forward, unet.py:1657-1686."""
    h = F.conv2d(F.silu(group_norm(sd, f"{name}.norm1", x, eps)), sd[f"{name}.conv1.weight"],
                 sd[f"{name}.conv1.bias"], padding=1)
    h = h + linear(sd, f"{name}.time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.conv2d(F.silu(group_norm(sd, f"{name}.norm2", h, eps)), sd[f"{name}.conv2.weight"],
                 sd[f"{name}.conv2.bias"], padding=1)
    if f"{name}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{name}.conv_shortcut.weight"], sd[f"{name}.conv_shortcut.bias"])
    return x + h


def temporal_resnet_block(sd: SD, name: str, x: torch.Tensor, temb: torch.Tensor, eps: float) -> torch.Tensor:
    """TemporalResnetBlock.forward, unet.py:1746-1769.  x [B,C,F,h,w], temb [B,F,1280]."""
    h = F.conv3d(F.silu(group_norm(sd, f"{name}.norm1", x, eps)), sd[f"{name}.conv1.weight"],
                 sd[f"{name}.conv1.bias"], padding=(1, 0, 0))
    te = linear(sd, f"{name}.time_emb_proj", F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
    h = h + te
    h = F.conv3d(F.silu(group_norm(sd, f"{name}.norm2", h, eps)), sd[f"{name}.conv2.weight"],
                 sd[f"{name}.conv2.bias"], padding=(1, 0, 0))
    return x + h


def res_block(sd: SD, r: ResSpec, x: torch.Tensor, temb: torch.Tensor, frames: int) -> torch.Tensor:
    """SpatioTemporalResBlock.forward, unet.py:1579-1608.  x [B*F,Cin,h,w], temb [B*F,1280]."""
    xs = resnet_block_2d(sd, f"{r.name}.spatial_res_block", x, temb, r.eps)
    bf, c, h, w = xs.shape
    b = bf // frames
    xs5 = xs.reshape(b, frames, c, h, w).permute(0, 2, 1, 3, 4)
    xt5 = temporal_resnet_block(sd, f"{r.name}.temporal_res_block", xs5,
                                temb.reshape(b, frames, -1), r.eps)
    y = alpha_blend(sd, f"{r.name}.time_mixer", xs5, xt5)
    return y.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def time_embedding(sd: SD, cfg: UNetConfig, timestep: torch.Tensor, added_time_ids: torch.Tensor) -> torch.Tensor:
    """unet.py:249-271: emb[B,1280] = MLP(sinusoid(t,320)) + MLP(sinusoid(ids.flatten(),256) -> [B,768])."""
    b = added_time_ids.shape[0]
    dt = sd["conv_in.weight"].dtype   # the model dtype the fp32 sinusoids are cast to (:263, :269); fp32 here = no-op
    t = timestep.reshape(-1).float().expand(b)
    emb = timestep_mlp(sd, "time_embedding", timestep_embedding(t, cfg.block_out_channels[0]).to(dt))
    te = timestep_embedding(added_time_ids.flatten().float(), cfg.addition_time_embed_dim).reshape(b, -1).to(dt)
    return emb + timestep_mlp(sd, "add_embedding", te)


def pose_net(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """PoseNet.forward, pose_net.py:40-48.  x [F,3,H,W] -> [F,320,H/8,W/8]."""
    from stableanimator_b200.arch import POSENET_CONVS
    for i, (_, _, _, stride, pad) in enumerate(POSENET_CONVS):
        x = F.silu(F.conv2d(x, sd[f"conv_layers.{2 * i}.weight"], sd[f"conv_layers.{2 * i}.bias"],
                            stride=stride, padding=pad))
    x = F.conv2d(x, sd["final_proj.weight"], sd["final_proj.bias"])
    return x * sd["scale"]


# ------------------------------------------------------------------------------------ whole UNet
@torch.no_grad()
def unet_forward(sd: SD, cfg: UNetConfig, sample: torch.Tensor, timestep: torch.Tensor,
                 encoder_hidden_states: torch.Tensor, added_time_ids: torch.Tensor,
                 pose_latents: Optional[torch.Tensor] = None, taps: Optional[dict] = None) -> torch.Tensor:
    """UNetSpatioTemporalConditionModel.forward, unet.py:231-362.

    sample [B,F,8,h,w]; timestep 0-d; encoder_hidden_states [B,5,1024]; added_time_ids [B,3];
    pose_latents [B*F,320,h,w] or None.  Returns [B,F,4,h,w].  The reference only runs B == 1
    (AlphaBlender broadcast, unet.py:455); alpha is a scalar, so B > 1 here means "B independent
    B=1 calls" and pose_latents may be given for a subset via zeros.
    ``taps``: optional dict filled with intermediate activations (for per-block parity tests).
    """
    b, frames = sample.shape[:2]
    emb = time_embedding(sd, cfg, timestep, added_time_ids)
    x = sample.flatten(0, 1)
    emb = emb.repeat_interleave(frames, dim=0)                                   # :278
    ehs = encoder_hidden_states.repeat_interleave(frames, dim=0)                 # :280
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)         # :283
    if pose_latents is not None:
        x = x + pose_latents                                                     # :284-286
    if taps is not None:
        taps["conv_in"] = x
    skips = [x]
    stages = build_stages(cfg)
    for st in stages:
        if st.kind == "down":                                                    # :1276-1305, :1201-1211
            for r, t in zip(st.res, st.tst):
                x = res_block(sd, r, x, emb, frames)
                if taps is not None:
                    taps[r.name] = x
                if t is not None:
                    x = transformer_spatio_temporal(sd, t, x, ehs, frames, cfg.num_tokens)
                    if taps is not None:
                        taps[t.name] = x
                skips.append(x)
            if st.downsample:
                x = F.conv2d(x, sd[f"{st.downsample}.weight"], sd[f"{st.downsample}.bias"],
                             stride=2, padding=1)                                # :1474
                skips.append(x)
        elif st.kind == "mid":                                                   # :1146-1166
            x = res_block(sd, st.res[0], x, emb, frames)
            x = transformer_spatio_temporal(sd, st.tst[0], x, ehs, frames, cfg.num_tokens)
            x = res_block(sd, st.res[1], x, emb, frames)
            if taps is not None:
                taps["mid_block"] = x
        else:                                                                    # :1336-1358, :1421-1456
            for r, t in zip(st.res, st.tst):
                x = torch.cat([x, skips.pop()], dim=1)
                x = res_block(sd, r, x, emb, frames)
                if t is not None:
                    x = transformer_spatio_temporal(sd, t, x, ehs, frames, cfg.num_tokens)
                if taps is not None:
                    taps[r.name] = x
            if st.upsample:
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")          # :1799
                x = F.conv2d(x, sd[f"{st.upsample}.weight"], sd[f"{st.upsample}.bias"], padding=1)
    x = F.silu(group_norm(sd, "conv_norm_out", x, 1e-5))                         # :355-356
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)       # :357
    return x.reshape(b, frames, *x.shape[1:])                                    # :360
