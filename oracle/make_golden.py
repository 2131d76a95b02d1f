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
