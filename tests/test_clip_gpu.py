"""GPU parity of the image-conditioning path (ClipImageEncoderB200) against the oracle restatement of
``_encode_image`` (inference_pipeline_animation.py:192-244) and the golden vectors produced by the reference's own
resize code + transformers' CLIPImageProcessor / CLIPVisionModelWithProjection."""
from pathlib import Path

import pytest
import torch

from util import assert_close, rel_l2

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _tiny():
    from stableanimator_b200.arch import ClipVisionConfig
    from stableanimator_b200.clip_encoder import ClipImageEncoderB200
    from stableanimator_b200.synth import synth_clip_vision_state_dict
    cfg = ClipVisionConfig.tiny()
    sd = synth_clip_vision_state_dict(cfg)
    return cfg, sd, ClipImageEncoderB200(cfg, sd, "cuda")


@pytest.mark.parametrize("h,w", [(128, 128), (144, 256), (512, 512), (1024, 576), (56, 56), (60, 300)])
def test_pixel_values_match_oracle(h, w):
    """Antialiased resize + CLIP normalisation: fp32 kernels vs the oracle (which equals the reference bit for bit)."""
    from oracle import clip_oracle
    from stableanimator_b200.synth import synth_image
    cfg, _, enc = _tiny()
    image01 = (synth_image(h, w, batch=2) + 1.0) / 2.0
    src = image01.cuda()
    keep = src.clone()
    out = enc.pixel_values(src)
    assert torch.equal(src, keep), "the caller's image must not be modified"
    ref = clip_oracle.clip_pixel_values(image01, cfg)
    assert out.dtype == torch.float32
    err = ((out.cpu() - ref).abs().max() / ref.abs().max()).item()
    assert err <= 2e-6, f"pixel_values {h}x{w}: max rel err {err:.2e}"   # fp32 summation-order differences only


def test_pixel_values_match_reference_golden():
    from stableanimator_b200.synth import synth_image
    g = torch.load(GOLD / "clip_vision_tiny.pt", map_location="cpu", weights_only=False)
    _, _, enc = _tiny()
    for name, c in g["cases"].items():
        image01 = (synth_image(c["h"], c["w"], g["seed"], batch=g["batch"]) + 1.0) / 2.0
        out = enc.pixel_values(image01.cuda()).cpu()
        assert ((out - c["pixel_values"]).abs().max() / c["pixel_values"].abs().max()).item() <= 2e-6, name


def test_attn_tokens_matches_sdpa():
    """sa_attn_tokens (257 tokens, head_dim 80, 16 heads) vs fp32 softmax attention on the same bf16 q|k|v."""
    import ctypes as C  # noqa: F401
    from stableanimator_b200 import ops
    from stableanimator_b200._lib import check, lib
    torch.manual_seed(3)
    for b, n, heads, d in [(1, 257, 16, 80), (2, 17, 4, 80), (1, 50, 3, 64), (1, 33, 2, 128)]:
        c = heads * d
        qkv = (torch.randn(b * n, 3 * c, device="cuda") * 1.5).bfloat16()
        out = torch.empty((b * n, c), device="cuda", dtype=torch.bfloat16)
        check(lib().sa_attn_tokens(qkv.data_ptr(), qkv.stride(0), b, n, heads, d, d ** -0.5, out.data_ptr(), out.stride(0),
                                   ops._stream()), "sa_attn_tokens")
        q, k, v = (t.float().view(b, n, heads, d).transpose(1, 2) for t in qkv.split(c, dim=1))
        ref = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v
        assert_close(out, ref.transpose(1, 2).reshape(b * n, c), what=f"attn_tokens b={b} n={n} h={heads} d={d}")


def test_image_embeds_match_oracle_and_golden():
    """Vision tower on the golden pixel values: vs transformers' image_embeds (fixture) and the fp32 oracle."""
    from oracle import clip_oracle
    g = torch.load(GOLD / "clip_vision_tiny.pt", map_location="cpu", weights_only=False)
    cfg, sd, enc = _tiny()
    for name, c in g["cases"].items():
        out = enc(c["pixel_values"].cuda()).image_embeds.cpu()
        ref = clip_oracle.clip_image_embeds(sd, cfg, c["pixel_values"])
        assert out.shape == ref.shape
        # fp32 output of a 2-layer bf16 tower: per-kernel tolerance chained over ~15 bf16 stages
        assert_close(out, ref, tol=6e-3, what=f"image_embeds vs oracle ({name})")
        assert_close(out, c["image_embeds"], tol=6e-3, what=f"image_embeds vs transformers ({name})")


def test_encode_image_cfg_layout_and_dtype():
    from oracle import clip_oracle
    from stableanimator_b200.synth import synth_image
    cfg, sd, enc = _tiny()
    image01 = (synth_image(96, 160) + 1.0) / 2.0
    out = enc.encode_image(image01.cuda())
    ref = clip_oracle.encode_image(sd, cfg, image01)
    assert out.shape == ref.shape == (2, 1, cfg.projection_dim)
    assert not out[0].any()                                     # unconditional half is zeros (:232-241)
    assert_close(out[1:].cpu(), ref[1:], tol=6e-3, what="encode_image")
    half = enc(enc.pixel_values(image01.cuda()).bfloat16()).image_embeds
    assert half.dtype == torch.bfloat16                          # the pipeline runs the encoder in the model dtype (:222)
    with pytest.raises(ValueError):
        enc.image_embeds(torch.zeros(1, 3, 32, 32, device="cuda"))
    bad = dict(sd)
    bad.pop("visual_projection.weight")
    from stableanimator_b200.clip_encoder import ClipImageEncoderB200
    with pytest.raises(KeyError):
        ClipImageEncoderB200(cfg, bad, "cuda")


def test_full_vit_h_matches_transformers_golden():
    """ViT-H/14 (32 layers, 1280 wide, 257 tokens) on a 512x512 image vs transformers' fp32 image_embeds (fixture).
    The bf16 residual stream through 64 sub-layers sets the floor; the same graph in torch bf16 is measured beside it."""
    from oracle import clip_oracle
    from stableanimator_b200.arch import ClipVisionConfig
    from stableanimator_b200.clip_encoder import ClipImageEncoderB200
    from stableanimator_b200.synth import synth_clip_vision_state_dict, synth_image
    g = torch.load(GOLD / "clip_vision_full.pt", map_location="cpu", weights_only=False)
    cfg = ClipVisionConfig.full()
    sd = synth_clip_vision_state_dict(cfg, g["seed"])
    enc = ClipImageEncoderB200(cfg, sd, "cuda")
    c = g["cases"]["a"]
    image01 = (synth_image(c["h"], c["w"], g["seed"], batch=g["batch"]) + 1.0) / 2.0
    pix = enc.pixel_values(image01.cuda())
    out = enc.image_embeds(pix)
    ref = c["image_embeds"].cuda()
    e = rel_l2(out, ref)
    sdc = {k: v.cuda().bfloat16() for k, v in sd.items()}
    floor = rel_l2(clip_oracle.clip_image_embeds(sdc, cfg, pix.bfloat16()).float(), ref)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        enc.image_embeds(pix)
    e1.record()
    torch.cuda.synchronize()
    print(f"ViT-H image_embeds rel-L2 vs transformers fp32: {e:.3e} (torch bf16 graph: {floor:.3e}); "
          f"{e0.elapsed_time(e1) / 3:.2f} ms per image")
    assert torch.isfinite(out).all()
    assert e <= 2e-2 and e <= max(2.0 * floor, 1e-2)
