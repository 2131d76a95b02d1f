"""CPU: checkpoint files (.pth and .safetensors) load into the packed layout exactly like the in-memory
state dict they were saved from; wrong files fail loudly (strict, as inference_basic.py:449-471)."""
import pytest
import torch

from stableanimator_b200 import checkpoint as ck
from stableanimator_b200.arch import UNetConfig, VaeDecoderConfig
from stableanimator_b200.synth import (synth_face_state_dict, synth_posenet_state_dict, synth_unet_state_dict,
                                       synth_vae_decoder_state_dict)


def _same(a, b, atol=0.0):
    assert a.keys() == b.keys()
    for k in a:
        if torch.is_tensor(a[k]):
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
            assert (a[k].double() - b[k].double()).abs().max().item() <= atol, k
        else:
            assert abs(a[k] - b[k]) <= atol, k


def test_unet_pth_and_safetensors_roundtrip(tmp_path):
    from safetensors.torch import save_file
    from stableanimator_b200.unet import UNetB200
    cfg = UNetConfig.tiny()
    sd = synth_unet_state_dict(cfg)
    ref = UNetB200(cfg, sd, "cpu")
    torch.save({k: v.half() for k, v in sd.items()}, tmp_path / "unet.pth")       # reference dtype: fp16
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "unet.safetensors"))
    a = ck.load_unet(tmp_path / "unet.pth", cfg, "cpu")
    b = ck.load_unet(tmp_path / "unet.safetensors", cfg, "cpu")
    _same(ref.w, b.w)
    # the fp16 file (the reference's weight dtype) packs to the same bf16 values except where fp16 goes
    # subnormal (|w| < 6.1e-5: absolute spacing 6e-8) or rounds a summed constant (rowsum, sigmoid(mix))
    _same(ref.w, a.w, atol=1e-3)
    for k in ("conv_in.w", "temb_all.w", "ctx_all.w"):
        assert (ref.w[k].double() - a.w[k].double()).abs().max().item() <= 6e-8, k


def test_strict_failures(tmp_path):
    cfg = UNetConfig.tiny()
    sd = synth_unet_state_dict(cfg)
    bad = dict(sd)
    bad.pop("conv_in.weight")
    torch.save(bad, tmp_path / "bad.pth")
    with pytest.raises(KeyError):
        ck.load_unet(tmp_path / "bad.pth", cfg, "cpu")
    with pytest.raises(FileNotFoundError):
        ck.load_unet(tmp_path / "nope.pth", cfg, "cpu")
    (tmp_path / "x.ckpt").write_bytes(b"")
    with pytest.raises(ValueError):
        ck.read_state_dict(tmp_path / "x.ckpt")


def test_other_modules_and_vae_prefix(tmp_path):
    from safetensors.torch import save_file
    torch.save(synth_posenet_state_dict(64), tmp_path / "pose_net.pth")
    torch.save(synth_face_state_dict(), tmp_path / "face_encoder.pth")
    assert ck.load_pose_net(tmp_path / "pose_net.pth", 64, "cpu").cout == 64
    assert ck.load_face_encoder(tmp_path / "face_encoder.pth", "cpu").c == 1024
    vcfg = VaeDecoderConfig.tiny()
    vsd = synth_vae_decoder_state_dict(vcfg)
    whole = {"decoder." + k: v.contiguous() for k, v in vsd.items()}
    whole["quant_conv.weight"] = torch.zeros(8, 8, 1, 1)      # the rest of the AutoencoderKLTemporalDecoder file
    whole["encoder.conv_in.weight"] = torch.zeros(64, 3, 3, 3)
    save_file(whole, str(tmp_path / "vae.safetensors"))
    dec = ck.load_vae_decoder(tmp_path / "vae.safetensors", vcfg, "cpu")
    assert dec.w["conv_in.w"].shape == (36, 128)


def test_image_encoder_checkpoint(tmp_path):
    """CLIPVisionModelWithProjection safetensors (incl. the legacy position_ids buffer) -> packed layout, strict."""
    from safetensors.torch import save_file
    from stableanimator_b200.arch import ClipVisionConfig
    from stableanimator_b200.synth import synth_clip_vision_state_dict
    cfg = ClipVisionConfig.tiny()
    sd = {k: v.contiguous() for k, v in synth_clip_vision_state_dict(cfg).items()}
    sd["vision_model.embeddings.position_ids"] = torch.arange(cfg.num_patches + 1)[None]
    save_file(sd, str(tmp_path / "model.safetensors"))
    enc = ck.load_image_encoder(tmp_path / "model.safetensors", cfg, "cpu")
    assert enc.w["patch.w"].shape == (cfg.hidden_size, 640) and enc.w["0.fc1.w"].shape == (2 * cfg.intermediate_size, cfg.hidden_size)
    assert enc.w["0.qkv.w"].shape == (3 * cfg.hidden_size, cfg.hidden_size)
    sd.pop("visual_projection.weight")
    save_file(sd, str(tmp_path / "bad.safetensors"))
    with pytest.raises(KeyError):
        ck.load_image_encoder(tmp_path / "bad.safetensors", cfg, "cpu")
