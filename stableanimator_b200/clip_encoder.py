"""B200-native image conditioning: ``InferenceAnimationPipeline._encode_image``
(animation/pipelines/inference_pipeline_animation.py:192-244) -- antialiased resize to the CLIP resolution,
``CLIPImageProcessor`` normalisation and the ``CLIPVisionModelWithProjection`` tower (third-party ``transformers``,
loaded at inference_basic.py:241-243) whose ``image_embeds`` become the CLIP token of ``encoder_hidden_states``.

Runs once per video.  Every arithmetic op is a ``libsa_b200.so`` call: ``sa_gauss_blur_axis`` +
``sa_resize_bicubic_norm`` (``_resize_with_antialiasing`` :99-125 fused with :210 and :213-220), ``sa_patchify`` +
``sa_gemm_tap`` (patch embedding, learned positions added in the epilogue), ``sa_layernorm``, ``sa_gemm_tap`` for
q|k|v (one fused projection), out_proj / fc2 (+ residual in the epilogue) and fc1 (exact GELU through the GEGLU
epilogue ``value * gelu(gate)`` with a constant-one value half), ``sa_attn_tokens`` (257 tokens, head_dim
80), and the fp32 skinny kernels for ``post_layernorm`` + ``visual_projection`` of the class token.
Same ``state_dict`` keys as the transformers module (strict); no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict

import torch

from . import ops
from ._lib import check, lib
from .arch import ClipVisionConfig, clip_vision_param_shapes


def _bf16(t, dev):
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class ClipImageEncoderB200:
    def __init__(self, cfg: ClipVisionConfig, state_dict: Dict[str, torch.Tensor], device="cuda"):
        self.cfg, self.dev = cfg, torch.device(device)
        sd = {k: v for k, v in state_dict.items() if not k.endswith("position_ids")}  # non-persistent buffer in old files
        shapes = clip_vision_param_shapes(cfg)
        missing = [k for k in shapes if k not in sd]
        extra = [k for k in sd if k not in shapes]
        if missing or extra:
            raise KeyError(f"CLIP vision state_dict mismatch: missing {missing[:5]} unexpected {extra[:5]}")
        for k, shp in shapes.items():
            if tuple(sd[k].shape) != tuple(shp):
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {shp}")
        if cfg.hidden_size % 64 or cfg.intermediate_size % 64 or (cfg.hidden_size // cfg.num_attention_heads) % 2:
            raise ValueError("hidden / intermediate sizes must be multiples of 64 and head_dim even")
        self.w: Dict[str, torch.Tensor] = {}
        self._pack(sd)

    def _pack(self, sd):
        w, dev, cfg, v = self.w, self.dev, self.cfg, "vision_model"
        c = cfg.hidden_size
        k = 3 * cfg.patch_size ** 2
        self.kpad = (k + 63) // 64 * 64
        wp = torch.zeros(c, self.kpad)
        wp[:, :k] = sd[f"{v}.embeddings.patch_embedding.weight"].float().reshape(c, k)
        w["patch.w"] = _bf16(wp, dev)
        pos = sd[f"{v}.embeddings.position_embedding.weight"].float()
        w["pos.patches"] = _f32(pos[1:], dev)                                     # epilogue row bias of the patch GEMM
        w["cls_row"] = _bf16(sd[f"{v}.embeddings.class_embedding"].float() + pos[0], dev)  # weight-only constant

        def norm(dst, src):
            w[f"{dst}.g"], w[f"{dst}.b"] = _f32(sd[f"{src}.weight"], dev), _f32(sd[f"{src}.bias"], dev)

        norm("pre", f"{v}.pre_layrnorm")
        norm("post", f"{v}.post_layernorm")
        for l in range(cfg.num_hidden_layers):
            p = f"{v}.encoder.layers.{l}"
            a = f"{p}.self_attn"
            w[f"{l}.qkv.w"] = _bf16(torch.cat([sd[f"{a}.q_proj.weight"], sd[f"{a}.k_proj.weight"],
                                               sd[f"{a}.v_proj.weight"]], 0), dev)
            w[f"{l}.qkv.b"] = _f32(torch.cat([sd[f"{a}.q_proj.bias"], sd[f"{a}.k_proj.bias"], sd[f"{a}.v_proj.bias"]], 0), dev)
            w[f"{l}.out.w"], w[f"{l}.out.b"] = _bf16(sd[f"{a}.out_proj.weight"], dev), _f32(sd[f"{a}.out_proj.bias"], dev)
            norm(f"{l}.ln1", f"{p}.layer_norm1")
            norm(f"{l}.ln2", f"{p}.layer_norm2")
            # gelu(fc1(x)) == value * gelu(gate) with value == 1 (zero weights, unit bias) and gate == fc1(x): the
            # GEGLU epilogue of sa_gemm_tap computes the exact-erf GELU, so no separate activation kernel exists
            w1, b1 = sd[f"{p}.mlp.fc1.weight"].float(), sd[f"{p}.mlp.fc1.bias"].float()
            wi, bi = ops.interleave_geglu(torch.cat([torch.zeros_like(w1), w1], 0), torch.cat([torch.ones_like(b1), b1], 0))
            w[f"{l}.fc1.w"], w[f"{l}.fc1.b"] = _bf16(wi, dev), _f32(bi, dev)
            w[f"{l}.fc2.w"], w[f"{l}.fc2.b"] = _bf16(sd[f"{p}.mlp.fc2.weight"], dev), _f32(sd[f"{p}.mlp.fc2.bias"], dev)
        w["proj.w"] = _bf16(sd["visual_projection.weight"], dev)

    # ------------------------------------------------------------------ preprocessing
    @staticmethod
    def _gauss_taps(k: int, sigma: float) -> torch.Tensor:
        """Filter coefficients of ``_gaussian`` (:67-80) -- a k-element host-side constant of the image size."""
        t = torch.arange(k, dtype=torch.float32) - k // 2
        g = torch.exp(-t.pow(2.0) / (2 * torch.tensor(sigma, dtype=torch.float32).pow(2.0)))
        return g / g.sum()

    @torch.no_grad()
    def pixel_values(self, image01: torch.Tensor) -> torch.Tensor:
        """``_encode_image`` :205-220: image [B, 3, H, W] in [0, 1] -> normalised fp32 [B, 3, S, S].
        ``image * 2 - 1`` (:208) is undone by ``(x + 1) / 2`` after the (linear, weights-sum-to-one) blur + bicubic
        resize, but both are kept so the arithmetic follows the reference step by step."""
        if image01.dim() != 4 or image01.shape[1] != 3:
            raise ValueError("image must be [B, 3, H, W]")
        cfg = self.cfg
        x = image01.to(self.dev, torch.float32).contiguous()
        b, _, h, wd = x.shape
        s = cfg.image_size
        sig, ks = [], []
        for n_in in (h, wd):
            sg = max((n_in / s - 1.0) / 2.0, 0.001)
            kk = int(max(2.0 * 2 * sg, 3))
            sig.append(sg)
            ks.append(kk + 1 if kk % 2 == 0 else kk)
        cur, bufs = x, [torch.empty_like(x), torch.empty_like(x)]   # x may be the caller's tensor: it is only read
        # columns first, then rows (:93-94); `image * 2.0 - 1.0` (:208) is applied to the values the first pass reads
        for axis, sg, kk, a0, c0 in ((1, sig[1], ks[1], 2.0, -1.0), (0, sig[0], ks[0], 1.0, 0.0)):
            taps = self._gauss_taps(kk, sg).to(self.dev)
            nxt = bufs.pop()
            check(lib().sa_gauss_blur_axis(cur.data_ptr(), nxt.data_ptr(), b * 3, h, wd, taps.data_ptr(), kk, axis,
                                           a0, c0, ops._stream()), "sa_gauss_blur_axis")
            cur = nxt
        out = torch.empty((b, 3, s, s), device=self.dev, dtype=torch.float32)
        mean = (C.c_float * 3)(*cfg.image_mean)
        std = (C.c_float * 3)(*cfg.image_std)
        check(lib().sa_resize_bicubic_norm(cur.data_ptr(), b, 3, h, wd, s, s, mean, std, out.data_ptr(), ops._stream()),
              "sa_resize_bicubic_norm")
        return out

    # ------------------------------------------------------------------ vision tower
    @torch.no_grad()
    def image_embeds(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """CLIPVisionModelWithProjection(pixel_values).image_embeds -> fp32 [B, projection_dim]."""
        cfg, w, dev = self.cfg, self.w, self.dev
        s = cfg.image_size
        if pixel_values.dim() != 4 or tuple(pixel_values.shape[1:]) != (3, s, s):
            raise ValueError(f"pixel_values must be [B, 3, {s}, {s}]")
        pix = pixel_values.to(dev, torch.float32).contiguous()
        b = pix.shape[0]
        c, heads, npatch = cfg.hidden_size, cfg.num_attention_heads, cfg.num_patches
        n, d = npatch + 1, cfg.hidden_size // cfg.num_attention_heads
        cols = torch.empty((b * npatch, self.kpad), device=dev, dtype=torch.bfloat16)
        check(lib().sa_patchify(pix.data_ptr(), b, 3, s, cfg.patch_size, cols.data_ptr(), self.kpad, ops._stream()),
              "sa_patchify")
        x = torch.empty((b * n, c), device=dev, dtype=torch.bfloat16)
        for i in range(b):
            x[i * n] = w["cls_row"]
            ops.gemm_tap(cols[i * npatch:(i + 1) * npatch], w["patch.w"], row_bias=w["pos.patches"], rows_per_group=1,
                         out=x[i * n + 1:(i + 1) * n])
        x = ops.layernorm(x, w["pre.g"], w["pre.b"], eps=cfg.layer_norm_eps)
        for l in range(cfg.num_hidden_layers):
            a = ops.layernorm(x, w[f"{l}.ln1.g"], w[f"{l}.ln1.b"], eps=cfg.layer_norm_eps)
            qkv = ops.gemm_tap(a, w[f"{l}.qkv.w"], bias=w[f"{l}.qkv.b"])
            o = torch.empty((b * n, c), device=dev, dtype=torch.bfloat16)
            check(lib().sa_attn_tokens(qkv.data_ptr(), qkv.stride(0), b, n, heads, d, float(d) ** -0.5, o.data_ptr(),
                                       o.stride(0), ops._stream()), "sa_attn_tokens")
            x = ops.gemm_tap(o, w[f"{l}.out.w"], bias=w[f"{l}.out.b"], residual=x)
            a = ops.layernorm(x, w[f"{l}.ln2.g"], w[f"{l}.ln2.b"], eps=cfg.layer_norm_eps)
            hmid = ops.gemm_tap(a, w[f"{l}.fc1.w"], bias=w[f"{l}.fc1.b"], geglu=True)
            x = ops.gemm_tap(hmid, w[f"{l}.fc2.w"], bias=w[f"{l}.fc2.b"], residual=x)
        cls = x.view(b, n, c)[:, 0].float().contiguous()                          # pooled = last_hidden_state[:, 0]
        pooled = torch.empty_like(cls)
        check(lib().sa_layernorm_f32(cls.data_ptr(), cls.stride(0), b, c, w["post.g"].data_ptr(), w["post.b"].data_ptr(),
                                     cfg.layer_norm_eps, pooled.data_ptr(), pooled.stride(0), ops._stream()),
              "sa_layernorm_f32")
        return ops.small_linear(pooled, w["proj.w"])

    def __call__(self, pixel_values: torch.Tensor):
        """Drop-in for ``self.image_encoder(image)`` (:224): an object with ``.image_embeds``."""
        emb = self.image_embeds(pixel_values)
        if pixel_values.dtype in (torch.bfloat16, torch.float16):
            emb = emb.to(pixel_values.dtype)
        return SimpleNamespace(image_embeds=emb)

    @torch.no_grad()
    def encode_image(self, image01: torch.Tensor, do_classifier_free_guidance: bool = True) -> torch.Tensor:
        """``_encode_image`` for one video per prompt (:192-244): [B, 3, H, W] in [0, 1] -> [2B | B, 1, proj], the
        zero unconditional embeddings first (:232-241)."""
        emb = self.image_embeds(self.pixel_values(image01)).unsqueeze(1)
        return torch.cat([torch.zeros_like(emb), emb]) if do_classifier_free_guidance else emb
