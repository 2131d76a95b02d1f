"""Checkpoint loading into the packed B200 layout.

The reference resumes three ``.pth`` state dicts with ``load_state_dict(strict=True)``
(inference_basic.py:449-471: ``checkpoints/Animation/{unet,pose_net,face_encoder}.pth``) and takes the VAE from
the SVD checkpoint (``vae/diffusion_pytorch_model.safetensors``, ``AutoencoderKLTemporalDecoder.from_pretrained``).
These helpers read the same files (``torch.load`` or safetensors) and hand the state dicts to the B200 modules,
whose constructors do the strict key / shape check and the packing (bf16 K-major matrices, fused q|k|v,
interleaved GEGLU rows, concatenated time / context projections).  ``dtype=torch.float16`` runs the UNet / PoseNet in
the reference's default inference dtype (inference_basic.py:486-503): fp16 checkpoints then keep their 11-bit
significands; with the default ``torch.bfloat16`` their matrices are re-rounded to bf16.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Union

import torch

from .arch import UNetConfig, VaeDecoderConfig

PathLike = Union[str, Path]


def read_state_dict(path: PathLike, prefix: str = "") -> Dict[str, torch.Tensor]:
    """``.pth`` / ``.pt`` / ``.bin`` via ``torch.load(map_location="cpu", weights_only=True)``,
    ``.safetensors`` via safetensors.  ``prefix`` keeps only the keys starting with it and strips it
    (``"decoder."`` selects the TemporalDecoder out of a whole-VAE file)."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(path)
    if path.suffix == ".safetensors":
        from safetensors.torch import load_file
        sd = load_file(str(path), device="cpu")
    elif path.suffix in (".pth", ".pt", ".bin"):
        sd = torch.load(str(path), map_location="cpu", weights_only=True)
        if isinstance(sd, dict) and "state_dict" in sd and all(not torch.is_tensor(v) for v in sd.values()):
            sd = sd["state_dict"]
    else:
        raise ValueError(f"unsupported checkpoint format: {path.suffix} (expected .pth/.pt/.bin/.safetensors)")
    if not isinstance(sd, dict) or not all(torch.is_tensor(v) for v in sd.values()):
        raise ValueError(f"{path} does not hold a flat state_dict of tensors")
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        if not sd:
            raise KeyError(f"{path}: no keys with prefix {prefix!r}")
    return sd


def load_unet(path: PathLike, cfg: Optional[UNetConfig] = None, device="cuda", dtype=torch.bfloat16):
    """``unet.pth`` (inference_basic.py:465-468) -> UNetB200."""
    from .unet import UNetB200
    return UNetB200(cfg or UNetConfig.full(), read_state_dict(path), device, dtype)


def load_pose_net(path: PathLike, noise_latent_channels: int = 320, device="cuda", dtype=torch.bfloat16):
    """``pose_net.pth`` (inference_basic.py:451-454) -> PoseNetB200."""
    from .pose_net import PoseNetB200
    return PoseNetB200(read_state_dict(path), noise_latent_channels, device, dtype)


def load_face_encoder(path: PathLike, device="cuda"):
    """``face_encoder.pth`` (inference_basic.py:458-461) -> FusionFaceIdB200."""
    from .face_encoder import FusionFaceIdB200
    return FusionFaceIdB200(read_state_dict(path), device)


def load_vae_decoder(path: PathLike, cfg: Optional[VaeDecoderConfig] = None, device="cuda"):
    """SVD ``vae/diffusion_pytorch_model.safetensors`` (whole AutoencoderKLTemporalDecoder: the ``decoder.``
    sub-tree is taken) or a bare TemporalDecoder state dict -> TemporalDecoderB200."""
    from .vae_decoder import TemporalDecoderB200
    sd = read_state_dict(path)
    if any(k.startswith("decoder.") for k in sd):
        sd = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    return TemporalDecoderB200(cfg or VaeDecoderConfig.full(), sd, device)


def load_vae_encoder(path: PathLike, cfg: Optional[VaeDecoderConfig] = None, device="cuda"):
    """SVD ``vae/diffusion_pytorch_model.safetensors`` -> VaeEncoderB200 (takes ``encoder.*`` and ``quant_conv.*``)."""
    from .vae_encoder import VaeEncoderB200
    return VaeEncoderB200(cfg or VaeDecoderConfig.full(), read_state_dict(path), device)


def load_image_encoder(path: PathLike, cfg=None, device="cuda"):
    """SVD ``image_encoder/model.safetensors`` (transformers CLIPVisionModelWithProjection, inference_basic.py:241-243)
    -> ClipImageEncoderB200."""
    from .arch import ClipVisionConfig
    from .clip_encoder import ClipImageEncoderB200
    return ClipImageEncoderB200(cfg or ClipVisionConfig.full(), read_state_dict(path), device)
