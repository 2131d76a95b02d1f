"""Stub of diffusers.utils.torch_utils for importing the reference vae.py (oracle/ref_loader.py)."""
import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)
