"""GPU parity of the temporal VAE decoder (TemporalDecoderB200) against the oracle restatement of
TemporalDecoder / decode_latents (vae.py:387-479, inference_pipeline_animation.py:319-357), against the
reference golden, and of its two helper kernels."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from util import assert_close, rel_l2

pytestmark = pytest.mark.gpu


def test_softmax_rows():
    from stableanimator_b200 import ops
    from stableanimator_b200._lib import check, lib
    m, n = 300, 1000
    s = torch.randn(m, n, device="cuda") * 20.0
    p = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    check(lib().sa_softmax_rows(s.data_ptr(), n, m, n, 0.125, p.data_ptr(), n, ops._stream()), "sa_softmax_rows")
    assert_close(p, torch.softmax(s * 0.125, dim=-1), what="softmax rows")


def test_time_conv_out():
    from stableanimator_b200 import ops
    from stableanimator_b200._lib import check, lib
    b, f, hw, c = 2, 4, 96, 3
    x = torch.randn(b * f * hw, 8, device="cuda")
    w = torch.randn(c, c, 3, device="cuda")
    bias = torch.randn(c, device="cuda")
    out = torch.empty(b * f, c, hw, device="cuda")
    check(lib().sa_time_conv_out(x.data_ptr(), 8, b, f, hw, c, w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                 ops._stream()), "sa_time_conv_out")
    x5 = x[:, :c].reshape(b, f, hw, c).permute(0, 3, 1, 2)[..., None]            # [b, c, f, hw, 1]
    ref = F.conv3d(x5, w[:, :, :, None, None], bias, padding=(1, 0, 0))[..., 0]     # [b, c, f, hw]
    ref = ref.permute(0, 2, 1, 3).reshape(b * f, c, hw)
    assert_close(out, ref, what="time_conv_out")


def test_vae_decoder_matches_oracle_and_reference_golden():
    from oracle import vae_oracle
    from stableanimator_b200.arch import VaeDecoderConfig
    from stableanimator_b200.synth import synth_vae_decoder_state_dict, synth_vae_latents
    from stableanimator_b200.vae_decoder import TemporalDecoderB200
    g = torch.load(Path(__file__).resolve().parent / "golden" / "vae_decoder_tiny.pt", map_location="cpu", weights_only=False)
    cfg = VaeDecoderConfig.tiny()
    sd = synth_vae_decoder_state_dict(cfg, g["seed"])
    lat = synth_vae_latents(g["frames"], g["h"], g["w"], g["seed"])
    dec = TemporalDecoderB200(cfg, sd, "cuda")
    out = dec.decode_latents(lat.cuda(), g["frames"], 4).cpu()
    ref = vae_oracle.decode_latents(sd, cfg, lat, g["frames"], 4)
    assert out.shape == ref.shape and out.dtype == torch.float32
    # Whole-decoder tolerance: rel-L2 <= 2.5e-2 against the fp32 path.  The decoder carries a bf16 residual
    # stream through 14 video residual blocks (56 GroupNorms) with random-init weights; the reference graph
    # itself, run in bf16 (torch CPU), is 1.88e-2 away from its fp32 result on this case -- that is the floor
    # of any bf16 implementation, ours measures 1.46e-2.  The reference golden is the fp32 result stored as fp16.
    l2 = rel_l2(out, ref)
    print(f"vae decoder rel-L2 vs fp32 oracle: {l2:.3e}; vs reference golden: {rel_l2(out, g['out'].float()):.3e}")
    assert l2 <= 2.5e-2
    assert rel_l2(out, g["out"].float()) <= 2.5e-2


def test_vae_decoder_forward_signature_and_strictness():
    from stableanimator_b200.arch import VaeDecoderConfig
    from stableanimator_b200.synth import synth_vae_decoder_state_dict
    from stableanimator_b200.vae_decoder import TemporalDecoderB200
    cfg = VaeDecoderConfig.tiny()
    sd = synth_vae_decoder_state_dict(cfg)
    bad = dict(sd)
    bad.pop("conv_in.bias")
    with pytest.raises(KeyError):
        TemporalDecoderB200(cfg, bad, "cuda")
    dec = TemporalDecoderB200(cfg, sd, "cuda")
    z = torch.randn(4, 4, 8, 8, device="cuda")
    y = dec(z, image_only_indicator=torch.zeros(1, 4, device="cuda"), num_frames=4)
    assert y.shape == (4, 3, 64, 64) and torch.isfinite(y).all()
    with pytest.raises(ValueError):
        dec(z, None, num_frames=3)


def test_im2col_s2_pad0_matches_asymmetric_pad_conv():
    """sa_im2col_s2_pad(pad=0) + GEMM == F.pad(x, (0,1,0,1)) + stride-2 conv (Downsample2D, vae.py:731-741)."""
    from stableanimator_b200 import ops
    n, h, w, c, co = 2, 12, 16, 64, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n, h, w, c, generator=g, device="cuda").bfloat16()
    wt = (torch.randn(co, c, 3, 3, generator=g, device="cuda") * (9 * c) ** -0.5).bfloat16()
    cols = ops.im2col_s2(x, pad=0)
    out = ops.gemm_tap(cols, wt.permute(0, 2, 3, 1).reshape(co, -1).contiguous())
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), wt.float(), stride=2)
    assert_close(out, ref.permute(0, 2, 3, 1).reshape(-1, co), what="downsample pad (0,1,0,1)")


def test_vae_encoder_matches_oracle_and_reference_golden():
    from oracle import vae_oracle
    from stableanimator_b200.arch import VaeDecoderConfig
    from stableanimator_b200.synth import synth_image, synth_vae_encoder_state_dict
    from stableanimator_b200.vae_encoder import VaeEncoderB200
    g = torch.load(Path(__file__).resolve().parent / "golden" / "vae_encoder_tiny.pt", map_location="cpu", weights_only=False)
    cfg = VaeDecoderConfig.tiny()
    sd = synth_vae_encoder_state_dict(cfg, g["seed"])
    img = synth_image(g["h"], g["w"], g["seed"], batch=g["batch"])
    enc = VaeEncoderB200(cfg, sd, "cuda")
    out = enc.encode_mode(img.cuda()).cpu()
    ref = vae_oracle.vae_encode_mode(sd, cfg, img)
    l2 = rel_l2(out, ref)
    print(f"vae encoder rel-L2 vs fp32 oracle: {l2:.3e}; vs reference golden: {rel_l2(out, g['out']):.3e}")
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert l2 <= 2.5e-2 and rel_l2(out, g["out"]) <= 2.5e-2   # bf16 activations through 11 residual blocks
    bad = dict(sd)
    bad.pop("quant_conv.bias")
    with pytest.raises(KeyError):
        VaeEncoderB200(cfg, bad, "cuda")
