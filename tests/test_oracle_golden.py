"""CPU: the oracle restatement reproduces the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py), and the scheduler restatement reproduces the constants the reference prints."""
import pytest
import torch

from cases import load_golden
from oracle import sampler_oracle as SO
from oracle import unet_oracle as O
from stableanimator_b200.synth import synth_posenet_state_dict, synth_unet_state_dict


@pytest.mark.parametrize("name", ["tiny_cond_f4_16x16", "tiny_uncond_f4_16x16", "tiny_cond_f2_8x24"])
def test_oracle_matches_reference_golden(name):
    c = load_golden(name)
    sd = synth_unet_state_dict(c["cfg"], c["gold"]["seed"])
    taps = {}
    out = O.unet_forward(sd, c["cfg"], c["sample"], c["t"], c["ehs"], c["ati"], c["pose_latents"], taps=taps)
    ref = c["gold"]["out"]
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 2e-5
    for k, v in c["gold"]["taps"].items():
        assert ((taps[k] - v).abs().max() / v.abs().max()).item() < 2e-5, k
    if c["pose_latents"] is not None:
        psd = synth_posenet_state_dict(c["cfg"].block_out_channels[0], c["gold"]["seed"])
        pl = O.pose_net(psd, c["pose"])
        assert ((pl - c["pose_latents"]).abs().max() / c["pose_latents"].abs().max()).item() < 1e-5


def test_scheduler_known_answers():
    """inference_pipeline_animation.py:405 (init_noise_sigma) and :633-639 (25 timesteps)."""
    sig = SO.karras_sigmas(25)
    ts = SO.timesteps_from_sigmas(sig)
    ref = torch.tensor(SO.REFERENCE_TIMESTEPS_25)
    assert (ts - ref).abs().max().item() < 2e-6
    assert abs(SO.init_noise_sigma(sig) - SO.REFERENCE_INIT_NOISE_SIGMA) < 1e-4
    assert float(sig[-1]) == 0.0 and abs(float(sig[0]) - 700.0) < 1e-3


def test_tile_indices_match_reference_recipe():
    """inference_pipeline_animation.py:613-616: slot 0 of every tile is frame 0; tail tile for F=128."""
    assert SO.tile_indices(16, 16, 4) == [list(range(16))]
    idx = SO.tile_indices(64, 16, 4)
    assert len(idx) == 5 and all(t[0] == 0 and len(t) == 16 for t in idx)
    assert [t[1] for t in idx] == [1, 13, 25, 37, 49] and idx[-1][-1] == 63
    idx = SO.tile_indices(128, 16, 4)
    assert len(idx) == 11 and idx[-1][1:] == list(range(113, 128)) and idx[-2][-1] == 123
    w = SO.tile_weights(16)
    assert abs(float(w[0]) - 0.0625) < 1e-7 and abs(float(w[7]) - 0.9375) < 1e-7 and torch.equal(w, w.flip(0))


def test_oracle_batch2_equals_two_b1_calls():
    c = load_golden("tiny_cond_f2_8x24")
    cfg = c["cfg"]
    sd = synth_unet_state_dict(cfg, c["gold"]["seed"])
    u = load_golden("tiny_uncond_f4_16x16")  # only for a second ehs/ati row (zeros)
    s2 = torch.cat([c["sample"] * 0.5, c["sample"]], 0)
    ehs2 = torch.cat([torch.zeros_like(c["ehs"]), c["ehs"]], 0)
    ati2 = torch.cat([c["ati"], c["ati"]], 0)
    pl = torch.cat([torch.zeros_like(c["pose_latents"]), c["pose_latents"]], 0)
    both = O.unet_forward(sd, cfg, s2, c["t"], ehs2, ati2, pl)
    a = O.unet_forward(sd, cfg, s2[:1], c["t"], ehs2[:1], ati2[:1], None)
    b = O.unet_forward(sd, cfg, s2[1:], c["t"], ehs2[1:], ati2[1:], c["pose_latents"])
    assert ((both - torch.cat([a, b], 0)).abs().max() / both.abs().max()).item() < 1e-5


def test_face_oracle_matches_reference_golden():
    """FusionFaceId (id_encoder.py:102-148): oracle/face_oracle.py vs the reference output stored by
    oracle/make_golden.py (same synthetic weights and embeddings, regenerated from the seed)."""
    from pathlib import Path
    from oracle import face_oracle
    from stableanimator_b200.synth import synth_face_inputs, synth_face_state_dict
    g = torch.load(Path(__file__).resolve().parent / "golden" / "face_fusion.pt", map_location="cpu", weights_only=False)
    sd = synth_face_state_dict(g["seed"])
    ide, clip = synth_face_inputs(g["batch"], g["seed"])
    out = face_oracle.fusion_face_id(sd, ide, clip)
    assert out.shape == g["out"].shape == (g["batch"], 4, 1024)
    assert ((out - g["out"]).abs().max() / g["out"].abs().max()).item() < 2e-5


def test_vae_oracle_matches_reference_golden():
    """TemporalDecoder + decode_latents (vae.py:387-479, inference_pipeline_animation.py:319-357): the oracle
    vs the reference frames stored (as fp16) by oracle/make_golden.py."""
    from pathlib import Path
    from oracle import vae_oracle
    from stableanimator_b200.arch import VaeDecoderConfig
    from stableanimator_b200.synth import synth_vae_decoder_state_dict, synth_vae_latents
    g = torch.load(Path(__file__).resolve().parent / "golden" / "vae_decoder_tiny.pt", map_location="cpu", weights_only=False)
    cfg = VaeDecoderConfig.tiny()
    sd = synth_vae_decoder_state_dict(cfg, g["seed"])
    lat = synth_vae_latents(g["frames"], g["h"], g["w"], g["seed"])
    out = vae_oracle.decode_latents(sd, cfg, lat, g["frames"], 4)
    ref = g["out"].float()
    assert out.shape == ref.shape == (1, 3, g["frames"], 8 * g["h"], 8 * g["w"])
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 1e-3  # fp16 storage of the fixture


def test_vae_encoder_oracle_matches_reference_golden():
    """Encoder + quant_conv + latent_dist.mode() (vae.py:236-256,482-556): oracle vs the stored reference output."""
    from pathlib import Path
    from oracle import vae_oracle
    from stableanimator_b200.arch import VaeDecoderConfig
    from stableanimator_b200.synth import synth_image, synth_vae_encoder_state_dict
    g = torch.load(Path(__file__).resolve().parent / "golden" / "vae_encoder_tiny.pt", map_location="cpu", weights_only=False)
    cfg = VaeDecoderConfig.tiny()
    sd = synth_vae_encoder_state_dict(cfg, g["seed"])
    out = vae_oracle.vae_encode_mode(sd, cfg, synth_image(g["h"], g["w"], g["seed"], batch=g["batch"]))
    assert out.shape == g["out"].shape == (g["batch"], 4, g["h"] // 8, g["w"] // 8)
    assert ((out - g["out"]).abs().max() / g["out"].abs().max()).item() < 2e-5


def test_clip_oracle_matches_reference_golden():
    """_encode_image (inference_pipeline_animation.py:192-244): oracle/clip_oracle.py vs the pixel values produced by
    the reference's own _resize_with_antialiasing + transformers' CLIPImageProcessor, and the image_embeds of
    transformers' CLIPVisionModelWithProjection on the same synthetic weights (stored by oracle/make_golden.py)."""
    from pathlib import Path
    from oracle import clip_oracle
    from stableanimator_b200.arch import ClipVisionConfig
    from stableanimator_b200.synth import synth_clip_vision_state_dict, synth_image
    g = torch.load(Path(__file__).resolve().parent / "golden" / "clip_vision_tiny.pt", map_location="cpu", weights_only=False)
    cfg = ClipVisionConfig.tiny()
    sd = synth_clip_vision_state_dict(cfg, g["seed"])
    for name, c in g["cases"].items():
        image01 = (synth_image(c["h"], c["w"], g["seed"], batch=g["batch"]) + 1.0) / 2.0
        pix = clip_oracle.clip_pixel_values(image01, cfg)
        assert ((pix - c["pixel_values"]).abs().max() / c["pixel_values"].abs().max()).item() < 1e-6, name
        emb = clip_oracle.clip_image_embeds(sd, cfg, pix)
        assert emb.shape == c["image_embeds"].shape == (g["batch"], cfg.projection_dim)
        assert ((emb - c["image_embeds"]).abs().max() / c["image_embeds"].abs().max()).item() < 2e-5, name
    both = clip_oracle.encode_image(sd, cfg, image01)
    assert both.shape == (2 * g["batch"], 1, cfg.projection_dim) and not both[:g["batch"]].any()


def test_clip_blur_geometry_of_the_baseline_resolutions():
    """Kernel sizes the reference derives (:100-120): 512x512 -> 3x3 taps; 1024x576 (h x w) -> 7 rows x 3 columns."""
    from oracle import clip_oracle
    assert clip_oracle.blur_geometry((512, 512), (224, 224))[1] == (3, 3)
    assert clip_oracle.blur_geometry((1024, 576), (224, 224))[1] == (7, 3)
    assert clip_oracle.blur_geometry((224, 224), (224, 224)) == ((0.001, 0.001), (3, 3))
    t = clip_oracle.gaussian_taps(7, 1.7857)
    assert abs(t.sum().item() - 1.0) < 1e-6 and torch.equal(t, t.flip(0))


def test_pipeline_oracle_chain_runs_and_is_deterministic():
    """oracle/pipeline_oracle.py (inference_pipeline_animation.py:494-706 as a chain of the pinned stage oracles):
    shapes, value ranges, the zero unconditional rows and determinism on the debug-size configs."""
    import dataclasses
    from oracle import pipeline_oracle
    from stableanimator_b200 import synth
    from stableanimator_b200.arch import ClipVisionConfig, UNetConfig, VaeDecoderConfig
    ucfg, vcfg = UNetConfig.tiny(), VaeDecoderConfig.tiny()
    ccfg = dataclasses.replace(ClipVisionConfig.tiny(), projection_dim=ucfg.cross_attention_dim)
    sd = dict(unet=synth.synth_unet_state_dict(ucfg), pose=synth.synth_posenet_state_dict(ucfg.block_out_channels[0]),
              face=synth.synth_face_state_dict(), clip=synth.synth_clip_vision_state_dict(ccfg),
              venc=synth.synth_vae_encoder_state_dict(vcfg), vdec=synth.synth_vae_decoder_state_dict(vcfg))
    frames, tile, H, W = 4, 4, 64, 64
    image01 = (synth.synth_image(H, W) + 1.0) / 2.0
    inp = synth.synth_inputs(ucfg, frames, H // 8, W // 8)
    ide = torch.randn(512, generator=torch.Generator().manual_seed(1))
    noise = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(2))
    run = lambda: pipeline_oracle.animation_pipeline(  # noqa: E731
        sd, (ucfg, vcfg, ccfg), image01, inp["pose"], ide, inp["latents"][:, :tile].contiguous(), noise,
        num_frames=frames, tile_size=tile, tile_overlap=1, num_inference_steps=2, decode_chunk_size=4)
    ehs, il, lat, video = run()
    assert ehs.shape == (2, 5, 1024) and not ehs[0].any() and ehs[1].abs().max() > 0
    assert il.shape == (2, frames, 4, H // 8, W // 8) and not il[0].any()
    assert lat.shape == (1, frames, 4, H // 8, W // 8) and torch.isfinite(lat).all()
    assert video.shape == (frames, 3, H, W) and video.min() >= 0 and video.max() <= 1
    assert torch.equal(run()[3], video)
