"""B200-native ``FusionFaceId`` (animation/modules/id_encoder.py:102-148): the face-token encoder of the
conditioning.  ``forward(id_embeds[B,512], clip_embeds[B,1,1024]) -> [B, 4, 1024]`` -- the four tokens the
pipeline appends to the CLIP token (inference_pipeline_animation.py:535-540) and the UNet's ID attention
reads (unet.py:1968-1972).  Runs once per video on <= 10 rows, so it is a chain of skinny kernels from
``libsa_b200.so`` with fp32 activations and bf16 weights: ``sa_small_linear`` (Linear, exact GELU),
``sa_layernorm_f32``, ``sa_perceiver_attn``.  Same ``state_dict`` keys as the reference module (strict)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import ops
from ._lib import check, lib
from .arch import face_param_shapes


class FusionFaceIdB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", num_tokens: int = 4, depth: int = 4,
                 heads: int = 16):
        self.dev = torch.device(device)
        self.num_tokens, self.depth, self.heads = num_tokens, depth, heads
        shapes = face_param_shapes(num_tokens=num_tokens, depth=depth)
        missing = [k for k in shapes if k not in state_dict]
        extra = [k for k in state_dict if k not in shapes]
        if missing or extra:
            raise KeyError(f"face encoder state_dict mismatch: missing {missing[:5]} unexpected {extra[:5]}")
        self.w: Dict[str, torch.Tensor] = {}
        for k, shp in shapes.items():
            t = state_dict[k]
            if tuple(t.shape) != tuple(shp):
                raise ValueError(f"{k}: shape {tuple(t.shape)} != {shp}")
            # matrices in bf16 (the tensor the skinny GEMV streams), vectors in fp32
            self.w[k] = t.detach().to(self.dev, torch.bfloat16 if t.dim() == 2 else torch.float32).contiguous()
        self.c = shapes["norm.weight"][0]

    def _ln(self, name: str, x: torch.Tensor) -> torch.Tensor:
        y = torch.empty_like(x)
        m, c = x.shape
        check(lib().sa_layernorm_f32(x.data_ptr(), x.stride(0), m, c, self.w[f"{name}.weight"].data_ptr(),
                                     self.w[f"{name}.bias"].data_ptr(), 1e-5, y.data_ptr(), y.stride(0),
                                     ops._stream()), "sa_layernorm_f32")
        return y

    def _lin(self, name: str, x: torch.Tensor, *, bias: bool = False, gelu: bool = False,
             accumulate_into: torch.Tensor = None) -> torch.Tensor:
        w = self.w[f"{name}.weight"]
        b = self.w[f"{name}.bias"] if bias else None
        out = accumulate_into if accumulate_into is not None else torch.empty((x.shape[0], w.shape[0]), device=self.dev)
        check(lib().sa_small_linear(x.data_ptr(), x.stride(0), x.shape[0], w.data_ptr(), w.shape[1], w.shape[0],
                                    None if b is None else b.data_ptr(), 0, 2 if gelu else 0,
                                    int(accumulate_into is not None), out.data_ptr(), out.stride(0), ops._stream()),
              "sa_small_linear")
        return out

    @torch.no_grad()
    def forward(self, id_embeds: torch.Tensor, clip_embeds: torch.Tensor, shortcut: bool = False,
                scale: float = 1.0) -> torch.Tensor:
        """Same arguments and asserts as FusionFaceId.forward (id_encoder.py:133-139)."""
        assert shortcut is False
        assert scale == 1.0
        if id_embeds.dim() != 2 or clip_embeds.dim() != 3 or clip_embeds.shape[0] != id_embeds.shape[0]:
            raise ValueError("id_embeds must be [B, 512] and clip_embeds [B, n, 1024]")
        bsz, c, nt = id_embeds.shape[0], self.c, self.num_tokens
        ide = id_embeds.to(self.dev, torch.float32).contiguous()
        clip = clip_embeds.to(self.dev, torch.float32).contiguous()
        nclip = clip.shape[1]
        h = self._lin("proj.0", ide, bias=True, gelu=True)
        x = self._lin("proj.2", h, bias=True).reshape(bsz * nt, c)
        latents = self._ln("norm", x)                                         # [B*4, C]
        xc = self._lin("fusion_model.proj_in", clip.reshape(bsz * nclip, -1), bias=True)   # [B*n, C]
        n_kv = nclip + nt
        for i in range(self.depth):
            a, f = f"fusion_model.layers.{i}.0", f"fusion_model.layers.{i}.1"
            # PerceiverAttention (id_encoder.py:33-57): keys/values over [norm1(x) ; norm2(latents)]
            ln_lat = self._ln(f"{a}.norm2", latents)
            kv_in = torch.empty((bsz, n_kv, c), device=self.dev)
            kv_in[:, :nclip] = self._ln(f"{a}.norm1", xc).reshape(bsz, nclip, c)
            kv_in[:, nclip:] = ln_lat.reshape(bsz, nt, c)
            q = self._lin(f"{a}.to_q", ln_lat)                                 # [B*4, C]
            kv = self._lin(f"{a}.to_kv", kv_in.reshape(bsz * n_kv, c))          # [B*(n+4), 2C]
            o = torch.empty((bsz * nt, c), device=self.dev)
            check(lib().sa_perceiver_attn(q.data_ptr(), q.stride(0), kv.data_ptr(), kv.stride(0), c, bsz, nt, n_kv,
                                          self.heads, o.data_ptr(), o.stride(0), ops._stream()), "sa_perceiver_attn")
            self._lin(f"{a}.to_out", o, accumulate_into=latents)               # latents += to_out(attn)
            hid = self._lin(f"{f}.1", self._ln(f"{f}.0", latents), gelu=True)   # [B*4, 4C]
            self._lin(f"{f}.3", hid, accumulate_into=latents)                  # latents += ff(latents)
        out = self._ln("fusion_model.norm_out", self._lin("fusion_model.proj_out", latents, bias=True))
        out = out.reshape(bsz, nt, c)
        return out.to(id_embeds.dtype) if id_embeds.dtype in (torch.bfloat16, torch.float16) else out

    __call__ = forward
