"""Architecture description of the StableAnimator / SVD spatio-temporal UNet.

Restates the module graph built by ``UNetSpatioTemporalConditionModel.__init__``
(animation/modules/unet.py:38-181) as plain data: the ordered block list, channel counts and the
reference ``state_dict`` key names + shapes (diffusers-SVD names, with the ID-attention projections
under ``...attn2.processor.id_to_{k,v}.weight`` as produced by ``set_attn_processor``,
unet.py:205-225 / inference_basic.py:348-377,445).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20, 20)
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 768
    cross_attention_dim: int = 1024
    layers_per_block: int = 2
    num_tokens: int = 4  # face tokens appended after the CLIP token(s)

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @staticmethod
    def full() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def tiny() -> "UNetConfig":
        """Same graph, 5x narrower (head_dim stays 64) -- for fast CPU oracle runs and GPU tests."""
        return UNetConfig(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4))


@dataclass
class ResSpec:
    """SpatioTemporalResBlock (unet.py:1546-1608)."""
    name: str
    cin: int
    cout: int
    eps: float
    skip: bool = False  # input is torch.cat([hidden, skip], dim=1) (unet.py:1347,1433)

    @property
    def shortcut(self) -> bool:
        return self.cin != self.cout


@dataclass
class TstSpec:
    """TransformerSpatioTemporalModel (unet.py:559-708)."""
    name: str
    c: int
    heads: int


@dataclass
class Stage:
    kind: str                 # "down" | "mid" | "up"
    name: str
    res: List[ResSpec] = field(default_factory=list)
    tst: List[Optional[TstSpec]] = field(default_factory=list)  # aligned with res ("mid": 1 entry)
    downsample: Optional[str] = None   # key prefix of the stride-2 conv
    upsample: Optional[str] = None     # key prefix of the post-nearest-x2 conv
    cout: int = 0


def build_stages(cfg: UNetConfig) -> List[Stage]:
    boc = cfg.block_out_channels
    heads = cfg.num_attention_heads
    stages: List[Stage] = []
    # ---- down (unet.py:97-121; eps: CrossAttnDown 1e-6 :1247, DownBlock 1e-5 :1191)
    out_c = boc[0]
    for i in range(4):
        in_c, out_c = out_c, boc[i]
        last = i == 3
        st = Stage("down", f"down_blocks.{i}", cout=out_c)
        for j in range(cfg.layers_per_block):
            cin = in_c if j == 0 else out_c
            st.res.append(ResSpec(f"down_blocks.{i}.resnets.{j}", cin, out_c, 1e-5 if last else 1e-6))
            st.tst.append(None if last else TstSpec(f"down_blocks.{i}.attentions.{j}", out_c, heads[i]))
        if not last:
            st.downsample = f"down_blocks.{i}.downsamplers.0.conv"
        stages.append(st)
    # ---- mid (unet.py:123-130, 1112-1140; eps 1e-5)
    c = boc[-1]
    mid = Stage("mid", "mid_block", cout=c)
    mid.res = [ResSpec("mid_block.resnets.0", c, c, 1e-5), ResSpec("mid_block.resnets.1", c, c, 1e-5)]
    mid.tst = [TstSpec("mid_block.attentions.0", c, heads[-1])]
    stages.append(mid)
    # ---- up (unet.py:132-168, 1320-1331, 1387-1407; eps = class default 1e-6)
    rev = list(reversed(boc))
    rheads = list(reversed(heads))
    out_c = rev[0]
    for i in range(4):
        prev_out = out_c
        out_c = rev[i]
        in_c = rev[min(i + 1, 3)]
        st = Stage("up", f"up_blocks.{i}", cout=out_c)
        nl = cfg.layers_per_block + 1
        for j in range(nl):
            skip_c = in_c if j == nl - 1 else out_c
            res_in = prev_out if j == 0 else out_c
            st.res.append(ResSpec(f"up_blocks.{i}.resnets.{j}", res_in + skip_c, out_c, 1e-6, skip=True))
            st.tst.append(None if i == 0 else TstSpec(f"up_blocks.{i}.attentions.{j}", out_c, rheads[i]))
        if i != 3:
            st.upsample = f"up_blocks.{i}.upsamplers.0.conv"
        stages.append(st)
    return stages


def _lin(d: Dict, name: str, cout: int, cin: int, bias: bool = True) -> None:
    d[f"{name}.weight"] = (cout, cin)
    if bias:
        d[f"{name}.bias"] = (cout,)


def _norm(d: Dict, name: str, c: int) -> None:
    d[f"{name}.weight"] = (c,)
    d[f"{name}.bias"] = (c,)


def _attn(d: Dict, name: str, c: int, ctx: int, id_proc: bool) -> None:
    _lin(d, f"{name}.to_q", c, c, bias=False)
    _lin(d, f"{name}.to_k", c, ctx, bias=False)
    _lin(d, f"{name}.to_v", c, ctx, bias=False)
    _lin(d, f"{name}.to_out.0", c, c)
    if id_proc:  # AnimationIDAttnNormalizedProcessor parameters (unet.py:1928-1929)
        _lin(d, f"{name}.processor.id_to_k", c, ctx, bias=False)
        _lin(d, f"{name}.processor.id_to_v", c, ctx, bias=False)


def _ff(d: Dict, name: str, c: int) -> None:
    _lin(d, f"{name}.net.0.proj", 8 * c, c)
    _lin(d, f"{name}.net.2", c, 4 * c)


def res_param_shapes(d: Dict, r: ResSpec, temb: int) -> None:
    s, t = f"{r.name}.spatial_res_block", f"{r.name}.temporal_res_block"
    _norm(d, f"{s}.norm1", r.cin)
    d[f"{s}.conv1.weight"] = (r.cout, r.cin, 3, 3)
    d[f"{s}.conv1.bias"] = (r.cout,)
    _lin(d, f"{s}.time_emb_proj", r.cout, temb)
    _norm(d, f"{s}.norm2", r.cout)
    d[f"{s}.conv2.weight"] = (r.cout, r.cout, 3, 3)
    d[f"{s}.conv2.bias"] = (r.cout,)
    if r.shortcut:
        d[f"{s}.conv_shortcut.weight"] = (r.cout, r.cin, 1, 1)
        d[f"{s}.conv_shortcut.bias"] = (r.cout,)
    _norm(d, f"{t}.norm1", r.cout)
    d[f"{t}.conv1.weight"] = (r.cout, r.cout, 3, 1, 1)
    d[f"{t}.conv1.bias"] = (r.cout,)
    _lin(d, f"{t}.time_emb_proj", r.cout, temb)
    _norm(d, f"{t}.norm2", r.cout)
    d[f"{t}.conv2.weight"] = (r.cout, r.cout, 3, 1, 1)
    d[f"{t}.conv2.bias"] = (r.cout,)
    d[f"{r.name}.time_mixer.mix_factor"] = (1,)


def tst_param_shapes(d: Dict, t: TstSpec, ctx: int) -> None:
    c = t.c
    _norm(d, f"{t.name}.norm", c)
    _lin(d, f"{t.name}.proj_in", c, c)
    b = f"{t.name}.transformer_blocks.0"
    _norm(d, f"{b}.norm1", c)
    _attn(d, f"{b}.attn1", c, c, False)
    _norm(d, f"{b}.norm2", c)
    _attn(d, f"{b}.attn2", c, ctx, True)
    _norm(d, f"{b}.norm3", c)
    _ff(d, f"{b}.ff", c)
    tb = f"{t.name}.temporal_transformer_blocks.0"
    _norm(d, f"{tb}.norm_in", c)
    _ff(d, f"{tb}.ff_in", c)
    _norm(d, f"{tb}.norm1", c)
    _attn(d, f"{tb}.attn1", c, c, False)
    _norm(d, f"{tb}.norm2", c)
    _attn(d, f"{tb}.attn2", c, ctx, False)
    _norm(d, f"{tb}.norm3", c)
    _ff(d, f"{tb}.ff", c)
    _lin(d, f"{t.name}.time_pos_embed.linear_1", 4 * c, c)
    _lin(d, f"{t.name}.time_pos_embed.linear_2", c, 4 * c)
    d[f"{t.name}.time_mixer.mix_factor"] = (1,)
    _lin(d, f"{t.name}.proj_out", c, c)


def unet_param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """Reference ``unet.state_dict()`` key -> shape (1460 entries for the full config)."""
    d: Dict[str, Tuple[int, ...]] = {}
    c0, temb = cfg.block_out_channels[0], cfg.time_embed_dim
    d["conv_in.weight"] = (c0, cfg.in_channels, 3, 3)
    d["conv_in.bias"] = (c0,)
    _lin(d, "time_embedding.linear_1", temb, c0)
    _lin(d, "time_embedding.linear_2", temb, temb)
    _lin(d, "add_embedding.linear_1", temb, cfg.projection_class_embeddings_input_dim)
    _lin(d, "add_embedding.linear_2", temb, temb)
    for st in build_stages(cfg):
        if st.kind == "mid":
            res_param_shapes(d, st.res[0], temb)
            tst_param_shapes(d, st.tst[0], cfg.cross_attention_dim)
            res_param_shapes(d, st.res[1], temb)
            continue
        for r, t in zip(st.res, st.tst):
            res_param_shapes(d, r, temb)
            if t is not None:
                tst_param_shapes(d, t, cfg.cross_attention_dim)
        for key in (st.downsample, st.upsample):
            if key:
                d[f"{key}.weight"] = (st.cout, st.cout, 3, 3)
                d[f"{key}.bias"] = (st.cout,)
    _norm(d, "conv_norm_out", c0)
    d["conv_out.weight"] = (cfg.out_channels, c0, 3, 3)
    d["conv_out.bias"] = (cfg.out_channels,)
    return d


def posenet_param_shapes(noise_latent_channels: int = 320) -> Dict[str, Tuple[int, ...]]:
    """PoseNet (animation/modules/pose_net.py:13-38): conv stack + 1x1 projection + scale."""
    spec = [(3, 3, 3), (3, 16, 4), (16, 16, 3), (16, 32, 4), (32, 32, 3), (32, 64, 4), (64, 64, 3),
            (64, 128, 3)]
    d: Dict[str, Tuple[int, ...]] = {}
    for i, (ci, co, k) in enumerate(spec):
        d[f"conv_layers.{2 * i}.weight"] = (co, ci, k, k)
        d[f"conv_layers.{2 * i}.bias"] = (co,)
    d["final_proj.weight"] = (noise_latent_channels, 128, 1, 1)
    d["final_proj.bias"] = (noise_latent_channels,)
    d["scale"] = (1,)
    return d


@dataclass
class VaeDecoderConfig:
    """TemporalDecoder of AutoencoderKLTemporalDecoder (animation/modules/vae.py:387-435)."""
    in_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    scaling_factor: float = 0.18215

    @staticmethod
    def full() -> "VaeDecoderConfig":
        return VaeDecoderConfig()

    @staticmethod
    def tiny() -> "VaeDecoderConfig":
        """Same graph, 4x narrower (channels stay multiples of 64 for the 64-channel TMA taps)."""
        return VaeDecoderConfig(block_out_channels=(64, 64, 128, 128))


def vae_decoder_blocks(cfg: VaeDecoderConfig):
    """[(block name, [(resnet name, cin, cout)], upsampler name | None)] in execution order
    (vae.py:395-416: mid block, then up blocks over the reversed channel list)."""
    rev = list(reversed(cfg.block_out_channels))
    c = rev[0]
    blocks = [("mid_block", [(f"mid_block.resnets.{i}", c, c) for i in range(cfg.layers_per_block)], None)]
    prev = c
    for i, co in enumerate(rev):
        res = [(f"up_blocks.{i}.resnets.{j}", prev if j == 0 else co, co) for j in range(cfg.layers_per_block + 1)]
        up = f"up_blocks.{i}.upsamplers.0.conv" if i < len(rev) - 1 else None
        blocks.append((f"up_blocks.{i}", res, up))
        prev = co
    return blocks


def vae_decoder_param_shapes(cfg: VaeDecoderConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys / shapes of the reference TemporalDecoder (vae.py:387-435,557-658,784-949)."""
    d: Dict[str, Tuple[int, ...]] = {}
    c_top = cfg.block_out_channels[-1]
    d["conv_in.weight"] = (c_top, cfg.in_channels, 3, 3)
    d["conv_in.bias"] = (c_top,)

    def norm(n, c):
        d[f"{n}.weight"] = (c,)
        d[f"{n}.bias"] = (c,)

    for _, res, up in vae_decoder_blocks(cfg):
        for name, ci, co in res:
            s_, t_ = f"{name}.spatial_res_block", f"{name}.temporal_res_block"
            norm(f"{s_}.norm1", ci)
            d[f"{s_}.conv1.weight"] = (co, ci, 3, 3); d[f"{s_}.conv1.bias"] = (co,)
            norm(f"{s_}.norm2", co)
            d[f"{s_}.conv2.weight"] = (co, co, 3, 3); d[f"{s_}.conv2.bias"] = (co,)
            if ci != co:
                d[f"{s_}.conv_shortcut.weight"] = (co, ci, 1, 1); d[f"{s_}.conv_shortcut.bias"] = (co,)
            norm(f"{t_}.norm1", co)
            d[f"{t_}.conv1.weight"] = (co, co, 3, 1, 1); d[f"{t_}.conv1.bias"] = (co,)
            norm(f"{t_}.norm2", co)
            d[f"{t_}.conv2.weight"] = (co, co, 3, 1, 1); d[f"{t_}.conv2.bias"] = (co,)
            d[f"{name}.time_mixer.mix_factor"] = (1,)
        if up:
            co = res[-1][2]
            d[f"{up}.weight"] = (co, co, 3, 3); d[f"{up}.bias"] = (co,)
    a = "mid_block.attentions.0"
    norm(f"{a}.group_norm", c_top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        d[f"{a}.{n}.weight"] = (c_top, c_top); d[f"{a}.{n}.bias"] = (c_top,)
    c0 = cfg.block_out_channels[0]
    norm("conv_norm_out", c0)
    d["conv_out.weight"] = (cfg.out_channels, c0, 3, 3); d["conv_out.bias"] = (cfg.out_channels,)
    d["time_conv_out.weight"] = (cfg.out_channels, cfg.out_channels, 3, 1, 1)
    d["time_conv_out.bias"] = (cfg.out_channels,)
    return d


def vae_encoder_blocks(cfg: VaeDecoderConfig):
    """[(resnet names with (cin, cout)), downsampler name | None] per DownEncoderBlock2D (vae.py:500-517)."""
    blocks, prev = [], cfg.block_out_channels[0]
    for i, co in enumerate(cfg.block_out_channels):
        res = [(f"down_blocks.{i}.resnets.{j}", prev if j == 0 else co, co) for j in range(cfg.layers_per_block)]
        down = f"down_blocks.{i}.downsamplers.0.conv" if i < len(cfg.block_out_channels) - 1 else None
        blocks.append((res, down))
        prev = co
    return blocks


def vae_encoder_param_shapes(cfg: VaeDecoderConfig, image_channels: int = 3) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys / shapes of ``Encoder`` (vae.py:482-556, keys prefixed ``encoder.``) plus ``quant_conv``
    (vae.py:247) of the reference AutoencoderKLTemporalDecoder."""
    d: Dict[str, Tuple[int, ...]] = {}
    c0, ct = cfg.block_out_channels[0], cfg.block_out_channels[-1]
    d["encoder.conv_in.weight"] = (c0, image_channels, 3, 3)
    d["encoder.conv_in.bias"] = (c0,)

    def resnet(n, ci, co):
        d[f"{n}.norm1.weight"] = (ci,); d[f"{n}.norm1.bias"] = (ci,)
        d[f"{n}.conv1.weight"] = (co, ci, 3, 3); d[f"{n}.conv1.bias"] = (co,)
        d[f"{n}.norm2.weight"] = (co,); d[f"{n}.norm2.bias"] = (co,)
        d[f"{n}.conv2.weight"] = (co, co, 3, 3); d[f"{n}.conv2.bias"] = (co,)
        if ci != co:
            d[f"{n}.conv_shortcut.weight"] = (co, ci, 1, 1); d[f"{n}.conv_shortcut.bias"] = (co,)

    for res, down in vae_encoder_blocks(cfg):
        for name, ci, co in res:
            resnet(f"encoder.{name}", ci, co)
        if down:
            co = res[-1][2]
            d[f"encoder.{down}.weight"] = (co, co, 3, 3); d[f"encoder.{down}.bias"] = (co,)
    for j in range(2):
        resnet(f"encoder.mid_block.resnets.{j}", ct, ct)
    a = "encoder.mid_block.attentions.0"
    d[f"{a}.group_norm.weight"] = (ct,); d[f"{a}.group_norm.bias"] = (ct,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        d[f"{a}.{n}.weight"] = (ct, ct); d[f"{a}.{n}.bias"] = (ct,)
    d["encoder.conv_norm_out.weight"] = (ct,); d["encoder.conv_norm_out.bias"] = (ct,)
    d["encoder.conv_out.weight"] = (2 * cfg.in_channels, ct, 3, 3); d["encoder.conv_out.bias"] = (2 * cfg.in_channels,)
    d["quant_conv.weight"] = (2 * cfg.in_channels, 2 * cfg.in_channels, 1, 1)
    d["quant_conv.bias"] = (2 * cfg.in_channels,)
    return d


def face_param_shapes(cross_attention_dim: int = 1024, id_embeddings_dim: int = 512, clip_embeddings_dim: int = 1024,
                      num_tokens: int = 4, depth: int = 4, ff_mult: int = 4) -> Dict[str, Tuple[int, ...]]:
    """FusionFaceId (animation/modules/id_encoder.py:102-131) with its FacePerceiver (:66-92)."""
    c, e = cross_attention_dim, id_embeddings_dim
    d: Dict[str, Tuple[int, ...]] = {
        "proj.0.weight": (2 * e, e), "proj.0.bias": (2 * e,),
        "proj.2.weight": (c * num_tokens, 2 * e), "proj.2.bias": (c * num_tokens,),
        "norm.weight": (c,), "norm.bias": (c,),
        "fusion_model.proj_in.weight": (c, clip_embeddings_dim), "fusion_model.proj_in.bias": (c,),
        "fusion_model.proj_out.weight": (c, c), "fusion_model.proj_out.bias": (c,),
        "fusion_model.norm_out.weight": (c,), "fusion_model.norm_out.bias": (c,),
    }
    for i in range(depth):
        a, f = f"fusion_model.layers.{i}.0", f"fusion_model.layers.{i}.1"
        d[f"{a}.norm1.weight"] = (c,); d[f"{a}.norm1.bias"] = (c,)
        d[f"{a}.norm2.weight"] = (c,); d[f"{a}.norm2.bias"] = (c,)
        d[f"{a}.to_q.weight"] = (c, c)
        d[f"{a}.to_kv.weight"] = (2 * c, c)
        d[f"{a}.to_out.weight"] = (c, c)
        d[f"{f}.0.weight"] = (c,); d[f"{f}.0.bias"] = (c,)
        d[f"{f}.1.weight"] = (ff_mult * c, c)
        d[f"{f}.3.weight"] = (c, ff_mult * c)
    return d


# PoseNet conv stack as (cin, cout, kernel, stride, pad)
POSENET_CONVS = [(3, 3, 3, 1, 1), (3, 16, 4, 2, 1), (16, 16, 3, 1, 1), (16, 32, 4, 2, 1),
                 (32, 32, 3, 1, 1), (32, 64, 4, 2, 1), (64, 64, 3, 1, 1), (64, 128, 3, 1, 1)]


@dataclass
class ClipVisionConfig:
    """``CLIPVisionModelWithProjection`` of the SVD ``image_encoder`` (third-party ``transformers``; loaded at
    inference_basic.py:241-243, called at inference_pipeline_animation.py:224): ViT-H/14, exact-erf GELU."""
    hidden_size: int = 1280
    intermediate_size: int = 5120
    num_hidden_layers: int = 32
    num_attention_heads: int = 16
    image_size: int = 224
    patch_size: int = 14
    projection_dim: int = 1024
    layer_norm_eps: float = 1e-5
    # CLIPImageProcessor statistics of the SVD ``feature_extractor`` (inference_basic.py:226)
    image_mean: Tuple[float, float, float] = (0.48145466, 0.4578275, 0.40821073)
    image_std: Tuple[float, float, float] = (0.26862954, 0.26130258, 0.27577711)

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @staticmethod
    def full() -> "ClipVisionConfig":
        return ClipVisionConfig()

    @staticmethod
    def tiny() -> "ClipVisionConfig":
        """Same graph: 2 layers, 4 heads of the real head_dim 80, 56x56 input (16 patches + class token)."""
        return ClipVisionConfig(hidden_size=320, intermediate_size=640, num_hidden_layers=2, num_attention_heads=4,
                                image_size=56, projection_dim=128)


def clip_vision_param_shapes(cfg: ClipVisionConfig) -> Dict[str, Tuple[int, ...]]:
    """``state_dict`` keys / shapes of transformers' CLIPVisionModelWithProjection (bias-free patch conv and
    projection; the misspelt ``pre_layrnorm`` is the library's own key)."""
    c, i, v = cfg.hidden_size, cfg.intermediate_size, "vision_model"
    d: Dict[str, Tuple[int, ...]] = {
        f"{v}.embeddings.class_embedding": (c,),
        f"{v}.embeddings.patch_embedding.weight": (c, 3, cfg.patch_size, cfg.patch_size),
        f"{v}.embeddings.position_embedding.weight": (cfg.num_patches + 1, c),
        f"{v}.pre_layrnorm.weight": (c,), f"{v}.pre_layrnorm.bias": (c,),
    }
    for l in range(cfg.num_hidden_layers):
        b = f"{v}.encoder.layers.{l}"
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            d[f"{b}.self_attn.{n}.weight"] = (c, c); d[f"{b}.self_attn.{n}.bias"] = (c,)
        d[f"{b}.layer_norm1.weight"] = (c,); d[f"{b}.layer_norm1.bias"] = (c,)
        d[f"{b}.mlp.fc1.weight"] = (i, c); d[f"{b}.mlp.fc1.bias"] = (i,)
        d[f"{b}.mlp.fc2.weight"] = (c, i); d[f"{b}.mlp.fc2.bias"] = (c,)
        d[f"{b}.layer_norm2.weight"] = (c,); d[f"{b}.layer_norm2.bias"] = (c,)
    d[f"{v}.post_layernorm.weight"] = (c,); d[f"{v}.post_layernorm.bias"] = (c,)
    d["visual_projection.weight"] = (cfg.projection_dim, c)
    return d
