"""GPU parity of the whole ``InferenceAnimationPipeline.__call__`` restated on the B200 modules (AnimationPipelineB200:
CLIP image encoder -> face tokens -> VAE encode -> denoising loop -> temporal VAE decode) against the chain of fp32
oracles, each of which is pinned to the unmodified reference / transformers (tests/test_oracle_golden.py)."""
import dataclasses

import pytest
import torch

from util import rel_l2

pytestmark = pytest.mark.gpu


def build(seed=23123134):
    from stableanimator_b200 import synth
    from stableanimator_b200.arch import ClipVisionConfig, UNetConfig, VaeDecoderConfig
    from stableanimator_b200.clip_encoder import ClipImageEncoderB200
    from stableanimator_b200.face_encoder import FusionFaceIdB200
    from stableanimator_b200.pipeline import AnimationPipelineB200
    from stableanimator_b200.pose_net import PoseNetB200
    from stableanimator_b200.unet import UNetB200
    from stableanimator_b200.vae_decoder import TemporalDecoderB200
    from stableanimator_b200.vae_encoder import VaeEncoderB200
    ucfg, vcfg = UNetConfig.tiny(), VaeDecoderConfig.tiny()
    ccfg = dataclasses.replace(ClipVisionConfig.tiny(), projection_dim=ucfg.cross_attention_dim)
    sd = dict(unet=synth.synth_unet_state_dict(ucfg, seed),
              pose=synth.synth_posenet_state_dict(ucfg.block_out_channels[0], seed),
              face=synth.synth_face_state_dict(seed), clip=synth.synth_clip_vision_state_dict(ccfg, seed),
              venc=synth.synth_vae_encoder_state_dict(vcfg, seed), vdec=synth.synth_vae_decoder_state_dict(vcfg, seed))
    pipe = AnimationPipelineB200(VaeEncoderB200(vcfg, sd["venc"], "cuda"), TemporalDecoderB200(vcfg, sd["vdec"], "cuda"),
                                 ClipImageEncoderB200(ccfg, sd["clip"], "cuda"), UNetB200(ucfg, sd["unet"], "cuda"),
                                 PoseNetB200(sd["pose"], ucfg.block_out_channels[0], "cuda"),
                                 FusionFaceIdB200(sd["face"], "cuda"))
    return pipe, sd, (ucfg, vcfg, ccfg)


def test_pipeline_matches_oracle_chain():
    from stableanimator_b200.synth import _gen, bf16_round, synth_image, synth_inputs
    pipe, sd, cfgs = build()
    frames, tile, overlap, steps, H, W = 6, 4, 2, 25, 128, 128
    image01 = (synth_image(H, W) + 1.0) / 2.0
    inp = synth_inputs(cfgs[0], frames, H // 8, W // 8)
    ide = bf16_round(torch.randn(512, generator=_gen("input.id_embeds", 1)))
    lat0 = inp["latents"][:, :tile].contiguous()
    img_noise = torch.randn(1, 3, H, W, generator=_gen("input.image_noise", 1))
    kw = dict(tile_size=tile, tile_overlap=overlap, num_inference_steps=steps, min_guidance_scale=1.0,
              max_guidance_scale=3.0, fps=7, motion_bucket_id=127, noise_aug_strength=0.02, decode_chunk_size=4)
    from oracle import pipeline_oracle
    ehs_r, il_r, lat_r, ref = pipeline_oracle.animation_pipeline(
        sd, cfgs, image01, inp["pose"], ide, lat0, img_noise, num_frames=frames, tile_size=tile, tile_overlap=overlap,
        num_inference_steps=steps, min_guidance_scale=1.0, max_guidance_scale=3.0, fps=7, motion_bucket_id=127,
        noise_aug_strength=0.02, decode_chunk_size=4)
    ehs = pipe.encode_conditioning(image01.cuda(), ide.numpy())
    assert ehs.shape == ehs_r.shape == (2, 5, 1024) and not ehs[0].any()
    e_ctx = rel_l2(ehs.cpu(), ehs_r)
    il = pipe.encode_image_latents(image01, H, W, frames, 0.02, img_noise)
    assert il.shape == il_r.shape and not il[0].any()
    e_il = rel_l2(il.cpu(), il_r)
    # the loop alone on the ORACLE's conditioning (identical inputs): the hot-path tolerance of north_star
    sampler = pipe._sampler(steps, tile, overlap, 1.0, 3.0)
    ati = torch.tensor([[6.0, 127.0, 0.02]] * 2)
    e_loop = rel_l2(sampler(lat0, il_r, ehs_r, ati, inp["pose"]).cpu(), lat_r)
    lat = pipe(image01, inp["pose"], height=H, width=W, num_frames=frames, latents=lat0, image_noise=img_noise,
               reference_image_id_ante_embedding=ide.numpy(), output_type="latent", **kw)
    e_lat = rel_l2(lat.cpu(), lat_r)
    out = pipe(image01, inp["pose"], height=H, width=W, num_frames=frames, latents=lat0, image_noise=img_noise,
               reference_image_id_ante_embedding=ide.numpy(), output_type="pt", **kw)
    assert len(out) == 1 and out[0].shape == (frames, 3, H, W)
    e_frames = rel_l2(out[0].cpu(), ref)
    print(f"pipeline vs oracle chain: conditioning {e_ctx:.2e}, image latents {e_il:.2e}, loop on oracle inputs "
          f"{e_loop:.2e}, end-to-end latents after {steps} steps {e_lat:.2e}, decoded frames {e_frames:.2e}")
    # per-stage bounds are the stages' own (tests/test_clip_gpu.py, test_vae_gpu.py, test_sampler_gpu.py; the loop on
    # identical inputs meets north_star's 1e-2 after 25 steps); end to end the bf16 VAE-encoder error of the conditioning
    # latents (1.7e-2) is an INPUT perturbation of every UNet call: bound the chain at 2e-2
    # (measured: 1.2e-3 / 1.7e-2 / 5.7e-3 / 7.4e-3 / 3.4e-3)
    assert e_ctx <= 1e-2 and e_il <= 2e-2 and e_loop <= 1e-2
    assert e_lat <= 2e-2 and e_frames <= 2e-2
    assert out[0].min() >= 0 and out[0].max() <= 1


def test_pipeline_input_forms_and_errors():
    """PIL reference image + uint8 pose frames (the reference's own input types), np / pil outputs, argument checks."""
    import numpy as np
    from PIL import Image
    pipe, _, cfgs = build()
    H = W = 64
    rng = np.random.default_rng(0)
    pil = Image.fromarray(rng.integers(0, 255, (80, 96, 3), dtype=np.uint8))          # resized to 64x64 by preprocess
    pose = [Image.fromarray(np.zeros((H, W, 3), dtype=np.uint8)) for _ in range(4)]
    ide = rng.standard_normal(512).astype(np.float32)
    g = torch.Generator().manual_seed(0)
    kw = dict(height=H, width=W, tile_size=4, tile_overlap=2, num_inference_steps=2, decode_chunk_size=4,
              reference_image_id_ante_embedding=ide)
    a = pipe(pil, pose, generator=g, output_type="np", **kw)
    assert a[0].shape == (4, H, W, 3) and np.isfinite(a[0]).all()
    b = pipe(pil, pose, generator=torch.Generator().manual_seed(0), output_type="pil", **kw)
    assert len(b[0]) == 4 and b[0][0].size == (W, H)
    assert np.abs(np.asarray(b[0][0], dtype=np.float32) / 255 - a[0][0]).max() <= 0.5 / 255 + 1e-6   # same seed, same video
    with pytest.raises(ValueError):
        pipe(pil, pose, **{**kw, "height": 60})
    with pytest.raises(ValueError):
        pipe(pil, pose, **{**kw, "reference_image_id_ante_embedding": None})
    with pytest.raises(ValueError):
        pipe(pil, pose[:3], num_frames=4, **kw)
