"""Import the UNMODIFIED reference modules from /root/reference through the stubs in oracle/stubs.

TEST INFRASTRUCTURE ONLY (build container; /root/reference does not exist on the GPU box).
Used by oracle/make_golden.py to pin oracle/unet_oracle.py against the real reference and to
generate the committed fixtures under tests/golden/.  Never imported by the product.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("SA_REFERENCE_ROOT", "/root/reference"))
STUBS = Path(__file__).resolve().parent / "stubs"


def available() -> bool:
    return (REFERENCE_ROOT / "animation" / "modules" / "unet.py").exists()


def import_reference():
    """Returns (unet_module, pose_net_module, id_encoder_module) of the reference."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for p in (str(STUBS), str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import animation.modules.unet as ref_unet  # noqa: E402
    import animation.modules.pose_net as ref_pose  # noqa: E402
    import animation.modules.id_encoder as ref_id  # noqa: E402
    return ref_unet, ref_pose, ref_id


def build_reference_unet(cfg, state_dict=None):
    """Construct the reference UNet for ``cfg`` (stableanimator_b200.arch.UNetConfig) and wire the
    64 attention processors exactly as inference_basic.py:348-377,445 does."""
    import torch
    ref_unet, _, _ = import_reference()
    unet = ref_unet.UNetSpatioTemporalConditionModel(
        in_channels=cfg.in_channels, out_channels=cfg.out_channels,
        block_out_channels=tuple(cfg.block_out_channels),
        addition_time_embed_dim=cfg.addition_time_embed_dim,
        projection_class_embeddings_input_dim=cfg.projection_class_embeddings_input_dim,
        layers_per_block=cfg.layers_per_block, cross_attention_dim=cfg.cross_attention_dim,
        transformer_layers_per_block=1, num_attention_heads=tuple(cfg.num_attention_heads))
    attn_procs = {}
    boc = cfg.block_out_channels
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            hidden = boc[-1]
        elif name.startswith("up_blocks"):
            hidden = list(reversed(boc))[int(name[len("up_blocks.")])]
        else:
            hidden = boc[int(name[len("down_blocks.")])]
        if "temporal_transformer_blocks" in name:
            attn_procs[name] = ref_unet.XFormersAttnProcessor()
        elif name.endswith("attn1.processor"):
            attn_procs[name] = ref_unet.AnimationAttnProcessor(hidden_size=hidden)
        else:
            attn_procs[name] = ref_unet.AnimationIDAttnNormalizedProcessor(
                hidden_size=hidden, cross_attention_dim=cfg.cross_attention_dim)
    unet.set_attn_processor(attn_procs)
    unet.requires_grad_(False).eval()
    if state_dict is not None:
        unet.load_state_dict({k: v.clone() for k, v in state_dict.items()}, strict=True)
    return unet.to(torch.float32)
