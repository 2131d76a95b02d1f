"""B200-native VAE image encoder: ``Encoder`` + ``quant_conv`` of ``AutoencoderKLTemporalDecoder``
(animation/modules/vae.py:236-256,482-556) as used by ``_encode_vae_image``
(animation/pipelines/inference_pipeline_animation.py:246-272) -- the conditioning image latents, once per
video (VAE half).

``encode_mode(image[B,3,H,W] in [-1,1]) -> [B,4,H/8,W/8]`` == ``vae.encode(image).latent_dist.mode()``.
Same ``state_dict`` keys as the reference VAE (``encoder.*``, ``quant_conv.*``; strict).  Kernels: 3x3 / 1x1
convolutions and attention projections on ``sa_gemm_tap``, GroupNorm(+SiLU) on ``sa_groupnorm``, conv_in
(3 channels) on ``sa_conv_direct``, the stride-2 ``Downsample2D`` (``F.pad(x, (0,1,0,1))`` + padding-0 conv,
vae.py:731-741) as ``sa_im2col_s2_pad`` + GEMM, the single-head mid-block attention as in the decoder, the
1x1 ``quant_conv`` on the fp32 conv_out result with ``sa_small_linear``.  The reference upcasts the VAE to fp32
for this call (force_upcast); here activations are bf16 with fp32 accumulation.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .arch import VaeDecoderConfig, vae_encoder_blocks, vae_encoder_param_shapes
from .vae_decoder import TemporalDecoderB200, _bf16, _f32


class VaeEncoderB200:
    def __init__(self, cfg: VaeDecoderConfig, state_dict: Dict[str, torch.Tensor], device="cuda"):
        self.cfg, self.dev = cfg, torch.device(device)
        if any(c % 64 for c in cfg.block_out_channels):
            raise ValueError("block_out_channels must be multiples of 64 (64-channel TMA taps)")
        shapes = vae_encoder_param_shapes(cfg)
        sd = {k: v for k, v in state_dict.items() if k.startswith("encoder.") or k.startswith("quant_conv.")}
        missing = [k for k in shapes if k not in sd]
        extra = [k for k in sd if k not in shapes]
        if missing or extra:  # strict over the encoder sub-tree (a whole-VAE state_dict may be passed)
            raise KeyError(f"vae encoder state_dict mismatch: missing {missing[:5]} unexpected {extra[:5]}")
        for k, shp in shapes.items():
            if tuple(sd[k].shape) != tuple(shp):
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {shp}")
        self.blocks = vae_encoder_blocks(cfg)
        self.w: Dict[str, torch.Tensor] = {}
        self._pack(sd)

    def _pack(self, sd):
        dev, w = self.dev, self.w

        def norm(n):
            w[f"{n}.g"], w[f"{n}.b"] = _f32(sd[f"{n}.weight"], dev), _f32(sd[f"{n}.bias"], dev)

        def conv3(n):
            t = sd[f"{n}.weight"]
            w[f"{n}.w"] = _bf16(t.permute(0, 2, 3, 1).reshape(t.shape[0], -1), dev)
            w[f"{n}.b"] = _f32(sd[f"{n}.bias"], dev)

        def resnet(n, ci, co):
            norm(f"{n}.norm1"); conv3(f"{n}.conv1"); norm(f"{n}.norm2"); conv3(f"{n}.conv2")
            if ci != co:
                t = sd[f"{n}.conv_shortcut.weight"]
                w[f"{n}.conv_shortcut.w"] = _bf16(t.reshape(t.shape[0], t.shape[1]), dev)
                w[f"{n}.conv_shortcut.b"] = _f32(sd[f"{n}.conv_shortcut.bias"], dev)

        t = sd["encoder.conv_in.weight"]
        w["encoder.conv_in.w"] = _bf16(t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]), dev)
        w["encoder.conv_in.b"] = _f32(sd["encoder.conv_in.bias"], dev)
        for res, down in self.blocks:
            for name, ci, co in res:
                resnet(f"encoder.{name}", ci, co)
            if down:
                conv3(f"encoder.{down}")
        ct = self.cfg.block_out_channels[-1]
        for j in range(2):
            resnet(f"encoder.mid_block.resnets.{j}", ct, ct)
        a = "encoder.mid_block.attentions.0"
        norm(f"{a}.group_norm")
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            w[f"{a}.{n}.w"], w[f"{a}.{n}.b"] = _bf16(sd[f"{a}.{n}.weight"], dev), _f32(sd[f"{a}.{n}.bias"], dev)
        norm("encoder.conv_norm_out")
        conv3("encoder.conv_out")  # N = 8 output columns: the fp32-output GEMM path
        q = sd["quant_conv.weight"]
        w["quant_conv.w"] = _bf16(q.reshape(q.shape[0], q.shape[1]), dev)
        w["quant_conv.b"] = _f32(sd["quant_conv.bias"], dev)

    def _resnet(self, name, ci, co, x, n, h, wd):
        """ResnetBlock2D.forward (vae.py:815-832) on tokens [n*h*w, C]."""
        w, S, grid = self.w, h * wd, (wd, h, n)
        n1 = ops.groupnorm(x, w[f"{name}.norm1.g"], w[f"{name}.norm1.b"], 1e-6, S, silu=True)
        h1 = ops.gemm_tap(n1, w[f"{name}.conv1.w"], grid=grid, taps=ops.TAPS_3X3, bias=w[f"{name}.conv1.b"])
        n2 = ops.groupnorm(h1, w[f"{name}.norm2.g"], w[f"{name}.norm2.b"], 1e-6, S, silu=True)
        sc = x if ci == co else ops.gemm_tap(x, w[f"{name}.conv_shortcut.w"], bias=w[f"{name}.conv_shortcut.b"])
        return ops.gemm_tap(n2, w[f"{name}.conv2.w"], grid=grid, taps=ops.TAPS_3X3, bias=w[f"{name}.conv2.b"], residual=sc)

    # the mid-block attention is the decoder's (vae.py:31-111): same kernels, this module's weights
    _attention = TemporalDecoderB200._attention

    @torch.no_grad()
    def encode_moments(self, image: torch.Tensor) -> torch.Tensor:
        """image [B, 3, H, W] -> posterior parameters fp32 [B, 8, H/8, W/8] (mean | logvar), vae.py:253-256."""
        cfg, w = self.cfg, self.w
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] % 8 or image.shape[3] % 8:
            raise ValueError("image must be [B, 3, H, W] with H and W multiples of 8")
        n, _, h, wd = image.shape
        s = image.to(self.dev)
        if s.dtype not in (torch.float32, torch.bfloat16):
            s = s.float()
        x_in = ops.nchw_to_nhwc(s.contiguous())
        c0 = cfg.block_out_channels[0]
        x = ops.conv_direct(x_in, w["encoder.conv_in.w"], w["encoder.conv_in.b"], c0, 3, 1, 1).reshape(-1, c0)
        for res, down in self.blocks:
            for name, ci, co in res:
                x = self._resnet(f"encoder.{name}", ci, co, x, n, h, wd)
            if down:
                co = res[-1][2]
                cols = ops.im2col_s2(x.reshape(n, h, wd, co), pad=0)
                x = ops.gemm_tap(cols, w[f"encoder.{down}.w"], bias=w[f"encoder.{down}.b"])
                h, wd = h // 2, wd // 2
        ct = cfg.block_out_channels[-1]
        x = self._resnet("encoder.mid_block.resnets.0", ct, ct, x, n, h, wd)
        x = self._attention(x, n, h * wd, a="encoder.mid_block.attentions.0")
        x = self._resnet("encoder.mid_block.resnets.1", ct, ct, x, n, h, wd)
        y = ops.groupnorm(x, w["encoder.conv_norm_out.g"], w["encoder.conv_norm_out.b"], 1e-6, h * wd, silu=True)
        m = ops.gemm_tap(y, w["encoder.conv_out.w"], grid=(wd, h, n), taps=ops.TAPS_3X3, bias=w["encoder.conv_out.b"],
                         out_f32=True)                                              # fp32 [n*h*w, 8]
        m = ops.small_linear(m, w["quant_conv.w"], w["quant_conv.b"])               # 1x1 quant_conv
        return ops.nhwc_to_nchw_f32(m, n, 2 * cfg.in_channels, h * wd).reshape(n, 2 * cfg.in_channels, h, wd)

    @torch.no_grad()
    def encode_mode(self, image: torch.Tensor) -> torch.Tensor:
        """``vae.encode(image).latent_dist.mode()`` (inference_pipeline_animation.py:257): [B, 4, H/8, W/8]."""
        return self.encode_moments(image)[:, : self.cfg.in_channels].contiguous()
