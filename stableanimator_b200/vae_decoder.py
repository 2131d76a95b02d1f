"""B200-native temporal VAE decoder: ``TemporalDecoder`` of ``AutoencoderKLTemporalDecoder``
(animation/modules/vae.py:387-479) and the pipeline's ``decode_latents``
(animation/pipelines/inference_pipeline_animation.py:319-357) -- the step right after the denoising loop
(SURVEY.md section 8f, N2).

Same ``state_dict`` as the reference module (strict).  Everything runs in ``libsa_b200.so`` on the UNet's
channel-innermost bf16 layout:
  * 3x3 / 1x1 / (3,1,1) convolutions and the attention projections: ``sa_gemm_tap`` (tcgen05 implicit GEMM);
    the video residual block folds the AlphaBlender (vae.py:957-965: ``(1 - s) x_s + s x_t`` with
    ``s = sigmoid(mix_factor)``, ``x_t = x_s + conv``) into the last GEMM's epilogue as ``x_s + s * conv``;
  * GroupNorm(+SiLU) per frame / per clip: ``sa_groupnorm`` (bulk-copy streamed);
  * mid-block attention (one head, head_dim = C, h*w tokens per frame, vae.py:60-111): per frame
    ``S = Q K^T`` (fp32 out), ``sa_softmax_rows``, ``O = P V`` with ``V^T`` produced directly by a GEMM with the
    roles of weights and activations swapped; the value bias is added after the product (rows of P sum to 1);
  * conv_in (4 input channels): ``sa_conv_direct``; conv_out (3 channels, fp32): ``sa_gemm_tap``;
    ``time_conv_out``: ``sa_time_conv_out``; nearest x2 + conv3x3: four phase convolutions (``ops.upsample_conv_phases``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import ops
from ._lib import check, lib
from .arch import VaeDecoderConfig, vae_decoder_blocks, vae_decoder_param_shapes


def _bf16(t, dev):
    return t.detach().to(dev, torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.detach().to(dev, torch.float32).contiguous()


class TemporalDecoderB200:
    def __init__(self, cfg: VaeDecoderConfig, state_dict: Dict[str, torch.Tensor], device="cuda"):
        self.cfg, self.dev = cfg, torch.device(device)
        if any(c % 64 for c in cfg.block_out_channels):
            raise ValueError("block_out_channels must be multiples of 64 (64-channel TMA taps)")
        shapes = vae_decoder_param_shapes(cfg)
        missing = [k for k in shapes if k not in state_dict]
        extra = [k for k in state_dict if k not in shapes]
        if missing or extra:  # strict=True semantics
            raise KeyError(f"vae decoder state_dict mismatch: missing {missing[:5]} unexpected {extra[:5]}")
        for k, shp in shapes.items():
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: shape {tuple(state_dict[k].shape)} != {shp}")
        self.blocks = vae_decoder_blocks(cfg)
        self.w: Dict[str, torch.Tensor] = {}
        self._pack(state_dict)

    def _pack(self, sd):
        dev, w = self.dev, self.w

        def norm(n):
            w[f"{n}.g"], w[f"{n}.b"] = _f32(sd[f"{n}.weight"], dev), _f32(sd[f"{n}.bias"], dev)

        def conv3(n):  # OIHW -> [O, (ky, kx, ci)]
            t = sd[f"{n}.weight"]
            w[f"{n}.w"] = _bf16(t.permute(0, 2, 3, 1).reshape(t.shape[0], -1), dev)
            w[f"{n}.b"] = _f32(sd[f"{n}.bias"], dev)

        def convt(n):  # OIDHW (3,1,1) -> [O, (kt, ci)]
            t = sd[f"{n}.weight"][:, :, :, 0, 0]
            w[f"{n}.w"] = _bf16(t.permute(0, 2, 1).reshape(t.shape[0], -1), dev)
            w[f"{n}.b"] = _f32(sd[f"{n}.bias"], dev)

        def lin(n):
            w[f"{n}.w"], w[f"{n}.b"] = _bf16(sd[f"{n}.weight"], dev), _f32(sd[f"{n}.bias"], dev)

        t = sd["conv_in.weight"]  # direct conv weights [k*k*cin, cout]
        w["conv_in.w"] = _bf16(t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]), dev)
        w["conv_in.b"] = _f32(sd["conv_in.bias"], dev)
        for _, res, up in self.blocks:
            for name, ci, co in res:
                s_, t_ = f"{name}.spatial_res_block", f"{name}.temporal_res_block"
                norm(f"{s_}.norm1"); conv3(f"{s_}.conv1"); norm(f"{s_}.norm2"); conv3(f"{s_}.conv2")
                if ci != co:
                    tt = sd[f"{s_}.conv_shortcut.weight"]
                    w[f"{s_}.conv_shortcut.w"] = _bf16(tt.reshape(tt.shape[0], tt.shape[1]), dev)
                    w[f"{s_}.conv_shortcut.b"] = _f32(sd[f"{s_}.conv_shortcut.bias"], dev)
                norm(f"{t_}.norm1"); convt(f"{t_}.conv1"); norm(f"{t_}.norm2"); convt(f"{t_}.conv2")
                # AlphaBlender (vae.py:957-965): weight of the temporal branch = sigmoid(mix_factor)
                w[f"{name}.s"] = float(torch.sigmoid(sd[f"{name}.time_mixer.mix_factor"].float()).item())
            if up:  # Upsample2D (vae.py: nearest x2 + conv3x3) as four 2x2-tap phase convolutions, see ops.upsample_conv_phases
                w[f"{up}.b"] = _f32(sd[f"{up}.bias"], dev)
                for i, (_, _, _, wp) in enumerate(ops.upsample_conv_phases(sd[f"{up}.weight"])):
                    w[f"{up}.ph{i}.w"] = _bf16(wp, dev)
        a = "mid_block.attentions.0"
        norm(f"{a}.group_norm")
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{a}.{n}")
        norm("conv_norm_out")
        t = sd["conv_out.weight"]  # pad N 3 -> 8 (zero rows) for the 8-column store granularity
        wo = torch.zeros(8, t.shape[1] * 9, dtype=torch.float32)
        wo[: t.shape[0]] = t.permute(0, 2, 3, 1).reshape(t.shape[0], -1).float().cpu()
        bo = torch.zeros(8)
        bo[: t.shape[0]] = sd["conv_out.bias"].float().cpu()
        w["conv_out.w"], w["conv_out.b"] = _bf16(wo, dev), _f32(bo, dev)
        w["time_conv_out.w"] = _f32(sd["time_conv_out.weight"][:, :, :, 0, 0], dev)  # [co, ci, kt]
        w["time_conv_out.b"] = _f32(sd["time_conv_out.bias"], dev)

    # ------------------------------------------------------------------ blocks (tokens [B*F*h*w, C])
    def _res_block(self, name, ci, co, x, B, F, h, wd):
        """SpatioTemporalResBlock.forward (vae.py:925-948)."""
        w = self.w
        S = h * wd
        s_, t_ = f"{name}.spatial_res_block", f"{name}.temporal_res_block"
        grid2d, gridt = (wd, h, B * F), (S, F, B)
        n1 = ops.groupnorm(x, w[f"{s_}.norm1.g"], w[f"{s_}.norm1.b"], 1e-6, S, silu=True)
        h1 = ops.gemm_tap(n1, w[f"{s_}.conv1.w"], grid=grid2d, taps=ops.TAPS_3X3, bias=w[f"{s_}.conv1.b"])
        n2 = ops.groupnorm(h1, w[f"{s_}.norm2.g"], w[f"{s_}.norm2.b"], 1e-6, S, silu=True)
        sc = x if ci == co else ops.gemm_tap(x, w[f"{s_}.conv_shortcut.w"], bias=w[f"{s_}.conv_shortcut.b"])
        xs = ops.gemm_tap(n2, w[f"{s_}.conv2.w"], grid=grid2d, taps=ops.TAPS_3X3, bias=w[f"{s_}.conv2.b"], residual=sc)
        # temporal branch: statistics over the whole clip (5-D GroupNorm, vae.py:880), 3-tap conv over frames
        n3 = ops.groupnorm(xs, w[f"{t_}.norm1.g"], w[f"{t_}.norm1.b"], 1e-5, F * S, silu=True)
        h2 = ops.gemm_tap(n3, w[f"{t_}.conv1.w"], grid=gridt, taps=ops.TAPS_T3, bias=w[f"{t_}.conv1.b"])
        n4 = ops.groupnorm(h2, w[f"{t_}.norm2.g"], w[f"{t_}.norm2.b"], 1e-5, F * S, silu=True)
        # (1 - s) xs + s (xs + conv2) == xs + s * conv2: one epilogue, one residual read
        return ops.gemm_tap(n4, w[f"{t_}.conv2.w"], grid=gridt, taps=ops.TAPS_T3, bias=w[f"{t_}.conv2.b"],
                            residual=xs, out_scale=w[f"{name}.s"])

    def _attention(self, x, BF, S, a="mid_block.attentions.0"):
        """Attention.forward (vae.py:60-111): one head of dimension C over the S tokens of each frame."""
        w = self.w
        c = x.shape[-1]
        n = ops.groupnorm(x, w[f"{a}.group_norm.g"], w[f"{a}.group_norm.b"], 1e-6, S, silu=False)
        q = ops.gemm_tap(n, w[f"{a}.to_q.w"], bias=w[f"{a}.to_q.b"])
        k = ops.gemm_tap(n, w[f"{a}.to_k.w"], bias=w[f"{a}.to_k.b"])
        o = torch.empty_like(x)
        sp = (S + 63) // 64 * 64  # the P V contraction runs over S keys: pad it to the 64-wide TMA tap with zeros
        scores = torch.empty((S, S), device=self.dev, dtype=torch.float32)
        probs = torch.zeros((S, sp), device=self.dev, dtype=torch.bfloat16)
        vt = torch.zeros((c, sp), device=self.dev, dtype=torch.bfloat16)
        scale = float(c) ** -0.5
        for f in range(BF):
            rows = slice(f * S, (f + 1) * S)
            # V^T [C, S] = W_v n_f^T: the weight matrix is the "activation", the frame's tokens the "weights"
            ops.gemm_tap(w[f"{a}.to_v.w"], n[rows], out=vt[:, :S])
            ops.gemm_tap(q[rows], k[rows], out=scores, out_f32=True)
            check(lib().sa_softmax_rows(scores.data_ptr(), S, S, S, scale, probs.data_ptr(), sp, ops._stream()),
                  "sa_softmax_rows")
            # rows of P sum to 1, so the value bias passes through the product unchanged
            ops.gemm_tap(probs, vt, bias=w[f"{a}.to_v.b"], out=o[rows])
        return ops.gemm_tap(o, w[f"{a}.to_out.0.w"], bias=w[f"{a}.to_out.0.b"], residual=x)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, sample: torch.Tensor, image_only_indicator=None, num_frames: int = 1) -> torch.Tensor:
        """Drop-in for TemporalDecoder.forward (vae.py:437-479): sample [B*F, 4, h, w] -> fp32 [B*F, 3, 8h, 8w].
        ``image_only_indicator`` is accepted for signature parity (the reference only takes its frame count)."""
        cfg, w = self.cfg, self.w
        if sample.dim() != 4 or sample.shape[1] != cfg.in_channels:
            raise ValueError(f"sample must be [B*F, {cfg.in_channels}, h, w]")
        BF, _, h, wd = sample.shape
        F = int(num_frames)
        if BF % F:
            raise ValueError("batch_frames must be a multiple of num_frames")
        if image_only_indicator is not None and image_only_indicator.shape[-1] != F:
            raise ValueError("image_only_indicator must be [B, num_frames]")
        B = BF // F
        s = sample.to(self.dev)
        if s.dtype not in (torch.float32, torch.bfloat16):
            s = s.float()
        x_in = ops.nchw_to_nhwc(s.contiguous())
        c_top = cfg.block_out_channels[-1]
        x = ops.conv_direct(x_in, w["conv_in.w"], w["conv_in.b"], c_top, 3, 1, 1).reshape(-1, c_top)
        for bname, res, up in self.blocks:
            for j, (rname, ci, co) in enumerate(res):
                if bname == "mid_block" and j >= 1:
                    x = self._attention(x, BF, h * wd)
                x = self._res_block(rname, ci, co, x, B, F, h, wd)
            if up:
                co = res[-1][2]
                xin = x.reshape(BF, h, wd, co)
                xu = torch.empty((BF, 2 * h, 2 * wd, co), device=self.dev, dtype=torch.bfloat16)
                for i, (py, px, ptaps) in enumerate(ops.UPSAMPLE_PHASE_TAPS):
                    ops.gemm_tap(xin, w[f"{up}.ph{i}.w"], grid=(wd, h, BF), taps=ptaps, bias=w[f"{up}.b"],
                                 out=xu[:, py::2, px::2, :])
                h, wd = 2 * h, 2 * wd
                x = xu.reshape(-1, co)
        n = ops.groupnorm(x, w["conv_norm_out.g"], w["conv_norm_out.b"], 1e-6, h * wd, silu=True)
        y = ops.gemm_tap(n, w["conv_out.w"], grid=(wd, h, BF), taps=ops.TAPS_3X3, bias=w["conv_out.b"], out_f32=True)
        out = torch.empty((BF, cfg.out_channels, h, wd), device=self.dev, dtype=torch.float32)
        check(lib().sa_time_conv_out(y.data_ptr(), y.stride(0), B, F, h * wd, cfg.out_channels,
                                     w["time_conv_out.w"].data_ptr(), w["time_conv_out.b"].data_ptr(),
                                     out.data_ptr(), ops._stream()), "sa_time_conv_out")
        return out

    __call__ = forward

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor, num_frames: int, decode_chunk_size: int = 4) -> torch.Tensor:
        """InferenceAnimationPipeline.decode_latents (inference_pipeline_animation.py:319-357):
        latents [B, F, 4, h, w] -> frames fp32 [B, 3, F, 8h, 8w], ``decode_chunk_size`` frames at a time."""
        z = latents.flatten(0, 1).to(self.dev).float() * (1.0 / self.cfg.scaling_factor)
        frames = []
        for i in range(0, z.shape[0], decode_chunk_size):
            chunk = z[i:i + decode_chunk_size]
            frames.append(self.forward(chunk, None, chunk.shape[0]))
        frames = torch.cat(frames, dim=0)
        return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()
