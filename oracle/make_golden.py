"""Generate tests/golden/*.pt from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

Imports /root/reference/animation/modules/{unet,pose_net}.py through oracle/stubs, loads the
deterministic synthetic weights (stableanimator_b200/synth.py) with ``strict=True``, runs the
reference in fp32 on synthetic inputs and stores inputs' recipe + outputs.  While doing so it asserts
that oracle/unet_oracle.py reproduces the reference to fp32 round-off (the oracle pin).
Fixtures are small (tiny config outputs and a few taps); weights are regenerated from the seed.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_loader, unet_oracle  # noqa: E402
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.synth import (synth_inputs, synth_posenet_state_dict,  # noqa: E402
                                       synth_unet_state_dict)

GOLD = ROOT / "tests" / "golden"
SEED = 23123134


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def unet_case(cfg: UNetConfig, name: str, frames: int, h: int, w: int, t: float, cond: bool):
    sd = synth_unet_state_dict(cfg, SEED)
    ref = ref_loader.build_reference_unet(cfg, sd)
    inp = synth_inputs(cfg, frames, h, w, SEED)
    row = 1 if cond else 0
    sample = torch.cat([inp["latents"].expand(1, frames, 4, h, w) * 3.0,
                        inp["image_latents"][row:row + 1]], dim=2)
    ehs = inp["encoder_hidden_states"][row:row + 1]
    ati = inp["added_time_ids"][row:row + 1]
    pose_latents = None
    if cond:
        _, rp, _ = ref_loader.import_reference()
        pn = rp.PoseNet(noise_latent_channels=cfg.block_out_channels[0]).eval()
        psd = synth_posenet_state_dict(cfg.block_out_channels[0], SEED)
        pn.load_state_dict(psd, strict=True)
        with torch.no_grad():
            pose_latents = pn(inp["pose"])
        mine_pose = unet_oracle.pose_net(psd, inp["pose"])
        e = rel(mine_pose, pose_latents)
        print(f"[{name}] pose_net oracle vs reference: {e:.2e}")
        assert e < 1e-5
    tt = torch.tensor(t)
    with torch.no_grad():
        out_ref = ref(sample, tt, ehs, ati, pose_latents=pose_latents)[0]
    taps = {}
    out_mine = unet_oracle.unet_forward(sd, cfg, sample, tt, ehs, ati, pose_latents, taps=taps)
    e = rel(out_mine, out_ref)
    print(f"[{name}] unet oracle vs reference: max rel err {e:.2e}; out std {out_ref.std():.4f}")
    assert e < 2e-5, "oracle restatement diverges from the reference"
    keep = {k: v.clone() for k, v in taps.items()
            if k in ("conv_in", "down_blocks.0.resnets.0", "down_blocks.0.attentions.0", "mid_block")}
    torch.save({"config": name, "frames": frames, "h": h, "w": w, "t": t, "cond": cond, "seed": SEED,
                "out": out_ref.clone(), "taps": keep,
                "pose_latents": None if pose_latents is None else pose_latents.clone()},
               GOLD / f"unet_{name}.pt")


def face_case():
    """FusionFaceId (id_encoder.py:102-148): pins oracle/face_oracle.py and writes tests/golden/face_fusion.pt."""
    from oracle import face_oracle
    from stableanimator_b200.synth import synth_face_inputs, synth_face_state_dict
    _, _, rid = ref_loader.import_reference()
    sd = synth_face_state_dict(SEED)
    m = rid.FusionFaceId().eval()
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    ide, clip = synth_face_inputs(2, SEED)
    with torch.no_grad():
        out_ref = m(ide, clip)
    out_mine = face_oracle.fusion_face_id(sd, ide, clip)
    e = rel(out_mine, out_ref)
    print(f"[face_fusion] oracle vs reference: {e:.2e}")
    assert e < 1e-5
    torch.save({"seed": SEED, "batch": 2, "out": out_ref.clone()}, GOLD / "face_fusion.pt")


def vae_case():
    """TemporalDecoder (vae.py:387-479) + decode_latents (inference_pipeline_animation.py:319-357): pins
    oracle/vae_oracle.py and writes tests/golden/vae_decoder_tiny.pt (8 frames = two chunks of 4, 8x12 latent)."""
    from oracle import vae_oracle
    from stableanimator_b200.arch import VaeDecoderConfig
    from stableanimator_b200.synth import synth_vae_decoder_state_dict, synth_vae_latents
    ref_loader.import_reference()
    import animation.modules.vae as ref_vae
    cfg = VaeDecoderConfig.tiny()
    sd = synth_vae_decoder_state_dict(cfg, SEED)
    dec = ref_vae.TemporalDecoder(in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                                  block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block).eval()
    dec.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    frames, h, w = 8, 8, 12
    lat = synth_vae_latents(frames, h, w, SEED)
    z = lat.flatten(0, 1) / cfg.scaling_factor
    outs = []
    with torch.no_grad():
        for i in range(0, frames, 4):  # decode_latents chunks of 4 (vae.decode asserts num_frames == 4)
            outs.append(dec(z[i:i + 4], image_only_indicator=torch.zeros(1, 4), num_frames=4))
    out_ref = torch.cat(outs, 0).reshape(1, frames, 3, 8 * h, 8 * w).permute(0, 2, 1, 3, 4)
    out_mine = vae_oracle.decode_latents(sd, cfg, lat, frames, 4)
    e = rel(out_mine, out_ref)
    print(f"[vae_decoder_tiny] oracle vs reference: {e:.2e}")
    assert e < 1e-5
    torch.save({"seed": SEED, "frames": frames, "h": h, "w": w, "out": out_ref.to(torch.float16).clone()},
               GOLD / "vae_decoder_tiny.pt")


def vae_encoder_case():
    """Encoder + quant_conv + latent_dist.mode() (vae.py:236-256,482-556; inference_pipeline_animation.py:257):
    pins oracle/vae_oracle.vae_encode_mode and writes tests/golden/vae_encoder_tiny.pt."""
    from oracle import vae_oracle
    from stableanimator_b200.arch import VaeDecoderConfig
    from stableanimator_b200.synth import synth_image, synth_vae_encoder_state_dict
    ref_loader.import_reference()
    import animation.modules.vae as ref_vae
    cfg = VaeDecoderConfig.tiny()
    sd = synth_vae_encoder_state_dict(cfg, SEED)
    vae = ref_vae.AutoencoderKLTemporalDecoder(block_out_channels=list(cfg.block_out_channels)).eval()
    missing, unexpected = vae.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    assert not unexpected and all(k.startswith("decoder.") for k in missing)
    img = synth_image(64, 96, SEED, batch=2)
    with torch.no_grad():
        out_ref = vae.encode(img).latent_dist.mode()
    out_mine = vae_oracle.vae_encode_mode(sd, cfg, img)
    e = rel(out_mine, out_ref)
    print(f"[vae_encoder_tiny] oracle vs reference: {e:.2e}")
    assert e < 1e-5
    torch.save({"seed": SEED, "h": 64, "w": 96, "batch": 2, "out": out_ref.clone()}, GOLD / "vae_encoder_tiny.pt")


def _reference_resize_fn():
    """The reference's own ``_resize_with_antialiasing`` and helpers, executed from the unmodified source file
    (inference_pipeline_animation.py:22-125).  The module itself cannot be imported offline (it pulls in diffusers'
    pipeline classes), so the five torch-only function definitions are compiled out of it with ``ast``."""
    import ast
    from typing import List
    src = (ref_loader.REFERENCE_ROOT / "animation/pipelines/inference_pipeline_animation.py").read_text()
    want = {"_compute_padding", "_filter2d", "_gaussian", "_gaussian_blur2d", "_resize_with_antialiasing"}
    tree = ast.parse(src)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in tree.body} == want
    ns = {"torch": torch, "List": List}
    exec(compile(tree, "inference_pipeline_animation.py", "exec"), ns)  # noqa: S102
    return ns["_resize_with_antialiasing"]


def clip_case(full: bool = False):
    """``_encode_image`` (inference_pipeline_animation.py:192-244): pins oracle/clip_oracle.py against the reference's
    own resize code and against transformers' CLIPImageProcessor / CLIPVisionModelWithProjection on the synthetic
    weights, and writes tests/golden/clip_vision_tiny.pt."""
    import transformers
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModelWithProjection
    from oracle import clip_oracle
    from stableanimator_b200.arch import ClipVisionConfig
    from stableanimator_b200.synth import synth_clip_vision_state_dict, synth_image
    cfg = ClipVisionConfig.full() if full else ClipVisionConfig.tiny()
    tag = "full" if full else "tiny"
    sd = synth_clip_vision_state_dict(cfg, SEED)
    m = CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
        num_attention_heads=cfg.num_attention_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
        projection_dim=cfg.projection_dim, hidden_act="gelu", layer_norm_eps=cfg.layer_norm_eps)).eval()
    missing, unexpected = m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    assert not unexpected and all(k.endswith("position_ids") for k in missing), (missing, unexpected)
    proc = CLIPImageProcessor(do_resize=False, do_center_crop=False, do_rescale=False, do_normalize=True,
                              image_mean=list(cfg.image_mean), image_std=list(cfg.image_std))
    resize = _reference_resize_fn()
    outs = {}
    # tiny: blur kernels (3, 3) and (3, 5); full: the 512x512 reference image of BASELINE configs[1] (ViT-H/14)
    sizes = {"a": (512, 512)} if full else {"a": (128, 128), "b": (144, 256)}
    batch = 1 if full else 2
    for name, (h, w) in sizes.items():
        image01 = (synth_image(h, w, SEED, batch=batch) + 1.0) / 2.0
        with torch.no_grad():
            r = resize(image01 * 2.0 - 1.0, (cfg.image_size, cfg.image_size))
            mine_r = clip_oracle.resize_with_antialiasing(image01 * 2.0 - 1.0, (cfg.image_size, cfg.image_size))
            e0 = rel(mine_r, r)
            pix = proc(images=(r + 1.0) / 2.0, return_tensors="pt").pixel_values
            mine_pix = clip_oracle.clip_pixel_values(image01, cfg)
            e1 = rel(mine_pix, pix)
            emb = m(pix).image_embeds
            mine = clip_oracle.clip_image_embeds(sd, cfg, mine_pix)
            e2 = rel(mine, emb)
        print(f"[clip_vision_{tag}/{name}] oracle vs reference resize {e0:.2e}, vs CLIPImageProcessor {e1:.2e}, "
              f"vs transformers {transformers.__version__} image_embeds {e2:.2e}")
        assert e0 < 1e-6 and e1 < 1e-5 and e2 < 1e-5
        outs[name] = {"h": h, "w": w, "image_embeds": emb.clone()}
        if not full:
            outs[name]["pixel_values"] = pix.clone()
    torch.save({"seed": SEED, "batch": batch, "cases": outs, "transformers": transformers.__version__},
               GOLD / f"clip_vision_{tag}.pt")


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    tiny = UNetConfig.tiny()
    unet_case(tiny, "tiny_cond_f4_16x16", 4, 16, 16, 1.637770, True)
    unet_case(tiny, "tiny_uncond_f4_16x16", 4, 16, 16, 0.338294, False)
    unet_case(tiny, "tiny_cond_f2_8x24", 2, 8, 24, -0.462234, True)
    if "--full" in sys.argv:
        unet_case(UNetConfig.full(), "full_cond_f2_8x8", 2, 8, 8, 1.064048, True)
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
    face_case()
    vae_case()
    vae_encoder_case()
    clip_case()
    if "--full" in sys.argv:
        clip_case(full=True)
