"""B200-native PoseNet (animation/modules/pose_net.py:10-48): 8 small convs + SiLU, 1x1 projection,
learnable output scale.  ``forward(x[F,3,H,W] in [-1,1]) -> [F, 320, H/8, W/8]``; the NHWC bf16
variant feeds the UNet's conv_in epilogue directly (unet.py:284-286).  It does not depend on the
timestep, so the sampler evaluates it once per frame tile instead of once per step
(inference_pipeline_animation.py:661 recomputes it 25x)."""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .arch import POSENET_CONVS, posenet_param_shapes


class PoseNetB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], noise_latent_channels: int = 320, device="cuda",
                 dtype=torch.bfloat16):
        if dtype not in ops.ACT_DTYPES:
            raise ValueError(f"dtype must be torch.bfloat16 or torch.float16, got {dtype}")
        self.dtype = dtype   # 16-bit format of activations and weights (fp16 = the reference's default, inference_basic.py:486)
        self.dev = torch.device(device)
        self.cout = noise_latent_channels
        shapes = posenet_param_shapes(noise_latent_channels)
        extra = [k for k in state_dict if k not in shapes]
        if extra:  # strict=True semantics of the reference loader (inference_basic.py:449-471)
            raise KeyError(f"pose_net state_dict: unexpected keys {extra[:5]}")
        for k, s in shapes.items():
            if k not in state_dict or tuple(state_dict[k].shape) != tuple(s):
                raise KeyError(f"pose_net state_dict: bad or missing {k}")
        self.layers = []
        for i, (ci, co, k, stride, pad) in enumerate(POSENET_CONVS):
            self.layers.append((self._pack(state_dict[f"conv_layers.{2 * i}.weight"]),
                                state_dict[f"conv_layers.{2 * i}.bias"].to(self.dev, torch.float32).contiguous(),
                                co, k, stride, pad))
        self.proj_w = self._pack(state_dict["final_proj.weight"])
        self.proj_b = state_dict["final_proj.bias"].to(self.dev, torch.float32).contiguous()
        self.scale = float(state_dict["scale"].float().item())

    def _pack(self, w: torch.Tensor) -> torch.Tensor:
        co, ci, k, _ = w.shape
        cpad = (co + 7) // 8 * 8
        p = torch.zeros(k * k * ci, cpad, dtype=torch.float32)
        p[:, :co] = w.detach().float().cpu().permute(2, 3, 1, 0).reshape(k * k * ci, co)
        return p.to(self.dev, self.dtype).contiguous()

    @torch.no_grad()
    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        """x [F,3,H,W] (fp32/bf16, NCHW) -> bf16 [F, H/8, W/8, 320] (NHWC)."""
        x = x.to(self.dev)
        if x.dtype not in (torch.float32, self.dtype):
            x = x.float()
        y = ops.nchw_to_nhwc(x.contiguous(), dtype=self.dtype)
        for wt, b, co, k, stride, pad in self.layers:
            y = ops.conv_direct(y, wt, b, co, k, stride, pad, silu=True)
        return ops.conv_direct(y, self.proj_w, self.proj_b, self.cout, 1, 1, 0, scale=self.scale)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference-shaped result: [F, 320, H/8, W/8] in x's dtype."""
        y = self.forward_nhwc(x)
        return y.permute(0, 3, 1, 2).to(x.dtype if x.dtype.is_floating_point else torch.float32)

    __call__ = forward
