"""fp16 activation mode: the reference runs ``.to('cuda', torch.float16)`` by default
(inference_basic.py:486-503).  Every kernel of the path takes the 16-bit format as a run-time flag (``act_f16``);
``UNetB200(..., dtype=torch.float16)`` / ``PoseNetB200(..., dtype=torch.float16)`` run the whole path in it.  Parity against
the same fp32 oracle and goldens as the bf16 mode, at the same tolerances (fp16 carries 11 significant bits instead of 8,
so the errors come out smaller), plus a checkpoint file round trip into a forward."""
import pytest
import torch
import torch.nn.functional as F

from cases import load_golden
from util import assert_close, rel_l2

pytestmark = pytest.mark.gpu
H = torch.float16


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(H)


def test_gemm_epilogues_fp16():
    from stableanimator_b200 import ops
    m, k, n = 2000, 320, 320
    a, w = rnd(m, k, seed=1), rnd(n, k, scale=k ** -0.5, seed=2)
    bias, res, bx = torch.randn(n, device="cuda"), rnd(m, n, seed=3), rnd(m, n, seed=4)
    out = ops.gemm_tap(a, w, bias=bias, residual=res, blend_x=bx, alpha=0.3, out_scale=0.5)
    assert out.dtype == H
    ref = 0.3 * bx.float() + 0.7 * ((a.float() @ w.float().t() + bias) * 0.5 + res.float())
    assert_close(out, ref, what="fp16 linear + residual + blend")
    assert rel_l2(out, ref) <= 6e-4                      # fp16 output rounding: ~2e-4
    wg, bg = rnd(8 * k, k, scale=k ** -0.5, seed=5), torch.randn(8 * k, device="cuda")
    wi, bi = ops.interleave_geglu(wg, bg)
    g = ops.gemm_tap(a, wi, bias=bi, geglu=True)
    val, gate = (a.float() @ wg.float().t() + bg).chunk(2, dim=-1)
    assert_close(g, val * F.gelu(gate), what="fp16 GEGLU")
    with pytest.raises(ValueError):
        ops.gemm_tap(a, w.bfloat16())                    # mixed formats are refused, not reinterpreted


def test_conv3x3_with_fused_groupnorm_statistics_fp16():
    from stableanimator_b200 import ops
    f, h, w_, cin, cout = 4, 16, 16, 128, 320
    x, wp = rnd(f, h, w_, cin, seed=6), rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5, seed=7)
    st = torch.empty((f, 32, 2), device="cuda", dtype=torch.float64)
    out = ops.gemm_tap(x, wp, grid=(w_, h, f), taps=ops.TAPS_3X3, gn_stats=st, gn_rows_per_stat=h * w_)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2), padding=1)
    assert_close(out, ref.permute(0, 2, 3, 1).reshape(-1, cout), what="fp16 conv3x3")
    v = out.double().reshape(f, h * w_, 32, cout // 32)
    assert torch.allclose(st[..., 0], v.sum(dim=(1, 3)), rtol=2e-6, atol=1e-4)
    gamma, beta = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    n = ops.groupnorm(out, gamma, beta, 1e-5, h * w_, silu=True, stats_ws=st, stats_ready=True)
    nref = F.silu(F.group_norm(out.float().reshape(f, h * w_, cout).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert_close(n, nref.reshape(-1, cout), what="fp16 GroupNorm+SiLU")


def test_norms_and_attention_fp16():
    from stableanimator_b200 import ops
    rows, c = 4096, 320
    x = rnd(rows, c, seed=8)
    g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    assert_close(ops.layernorm(x, g, b), F.layer_norm(x.float(), (c,), g, b), what="fp16 LayerNorm")
    big = rnd(32 * 4096, c, seed=9)                       # streamed kernels (>= 8 MB)
    assert_close(ops.layernorm(big, g, b), F.layer_norm(big.float(), (c,), g, b), what="fp16 LayerNorm (streamed)")
    n = ops.groupnorm(big, g, b, 1e-5, 4096, silu=True)
    nref = F.silu(F.group_norm(big.float().reshape(32, 4096, c).permute(0, 2, 1), 32, g, b, 1e-5)).permute(0, 2, 1)
    assert_close(n, nref.reshape(-1, c), what="fp16 GroupNorm+SiLU (streamed)")
    # spatial attention (tcgen05: q, k, v all fp16), temporal attention (mma.sync f16)
    frames, s, heads = 2, 1024, 2
    cc = heads * 64
    qkv = rnd(frames * s, 3 * cc, seed=10)
    q, k, v = (t.float().reshape(frames, s, heads, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(frames * s, cc)
    assert_close(ops.attn_spatial(qkv, frames, s, heads), ref, what="fp16 spatial attention")
    bsz, fr, sp = 2, 16, 64
    qkv = rnd(bsz * fr * sp, 3 * cc, seed=11)
    t5 = qkv.float().reshape(bsz, fr, sp, 3, heads, 64)
    q, k, v = (t5[:, :, :, i].permute(0, 2, 3, 1, 4) for i in range(3))            # [b, s, h, f, d]
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 3, 1, 2, 4).reshape(bsz * fr * sp, cc)
    assert_close(ops.attn_temporal(qkv, bsz, fr, sp, heads), ref, what="fp16 temporal attention")


def test_ff_fused_fp16():
    from stableanimator_b200 import ops
    c, m = 320, 4096
    x = rnd(m, c, seed=12)
    w1, b1 = rnd(8 * c, c, scale=c ** -0.5, seed=13), torch.randn(8 * c, device="cuda") * 0.5
    w2, b2 = rnd(c, 4 * c, scale=(4 * c) ** -0.5, seed=14), torch.randn(c, device="cuda") * 0.5
    g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
    wi, bi = ops.interleave_geglu(w1, b1)
    out = ops.ff_fused(x, g, b, wi, bi, w2, b2)
    a = F.layer_norm(x.float(), (c,), g, b).to(H).float()
    val, gate = (a @ w1.float().t() + b1).chunk(2, dim=-1)
    ref = (val * F.gelu(gate)).to(H).float() @ w2.float().t() + b2 + x.float()
    assert out.dtype == H
    assert_close(out, ref, what="fp16 fused feed-forward")


@pytest.mark.parametrize("name", ["tiny_cond_f4_16x16", "tiny_uncond_f4_16x16", "full_cond_f2_8x8"])
def test_unet_fp16_matches_reference_golden(name):
    from stableanimator_b200.pose_net import PoseNetB200
    from stableanimator_b200.synth import synth_posenet_state_dict, synth_unet_state_dict
    from stableanimator_b200.unet import UNetB200
    c = load_golden(name)
    m = UNetB200(c["cfg"], synth_unet_state_dict(c["cfg"], c["gold"]["seed"]), "cuda", dtype=H)
    pl = None
    if c["pose_latents"] is not None:
        c0 = c["cfg"].block_out_channels[0]
        pl = PoseNetB200(synth_posenet_state_dict(c0, c["gold"]["seed"]), c0, "cuda", dtype=H)(c["pose"].cuda())
    out = m(c["sample"].cuda(), c["t"], c["ehs"].cuda(), c["ati"].cuda(), pose_latents=pl)[0]
    e = rel_l2(out, c["gold"]["out"].cuda())
    print(f"{name}: fp16-mode UNet vs reference fp32 golden rel-L2 {e:.3e}")
    assert torch.isfinite(out).all() and e <= 1e-2
    # fp16 inputs in -> fp16 out, as the reference pipeline runs it
    out16 = m(c["sample"].cuda().half(), c["t"], c["ehs"].cuda().half(), c["ati"].cuda().half(), pose_latents=pl)[0]
    assert out16.dtype == H and rel_l2(out16, c["gold"]["out"].cuda()) <= 1e-2


def test_checkpoint_files_round_trip_into_an_fp16_forward(tmp_path):
    """unet.pth / pose_net.pth as the reference saves them (fp16 tensors, inference_basic.py:449-471) -> checkpoint.py ->
    fp16-mode forward == the model built from the in-memory state dict, bit for bit; safetensors likewise."""
    from safetensors.torch import save_file
    from stableanimator_b200 import checkpoint as ck
    from stableanimator_b200.pose_net import PoseNetB200
    from stableanimator_b200.synth import synth_posenet_state_dict, synth_unet_state_dict
    from stableanimator_b200.unet import UNetB200
    c = load_golden("tiny_cond_f4_16x16")
    cfg = c["cfg"]
    sd = {k: v.half().float() for k, v in synth_unet_state_dict(cfg, c["gold"]["seed"]).items()}   # fp16-representable
    psd = {k: v.half().float() for k, v in synth_posenet_state_dict(cfg.block_out_channels[0], c["gold"]["seed"]).items()}
    torch.save({k: v.half() for k, v in sd.items()}, tmp_path / "unet.pth")
    torch.save({k: v.half() for k, v in psd.items()}, tmp_path / "pose_net.pth")
    save_file({k: v.half().contiguous() for k, v in sd.items()}, str(tmp_path / "unet.safetensors"))
    args = (c["sample"].cuda(), c["t"], c["ehs"].cuda(), c["ati"].cuda())
    pose_ref = PoseNetB200(psd, cfg.block_out_channels[0], "cuda", dtype=H)(c["pose"].cuda())
    ref = UNetB200(cfg, sd, "cuda", dtype=H)(*args, pose_latents=pose_ref)[0]
    pose = ck.load_pose_net(tmp_path / "pose_net.pth", cfg.block_out_channels[0], "cuda", dtype=H)(c["pose"].cuda())
    assert torch.equal(pose, pose_ref)
    for f in ("unet.pth", "unet.safetensors"):
        out = ck.load_unet(tmp_path / f, cfg, "cuda", dtype=H)(*args, pose_latents=pose)[0]
        assert torch.equal(out, ref), f
    assert rel_l2(ref, c["gold"]["out"].cuda()) <= 1e-2


def test_sampler_fp16_25_steps_tiny():
    from oracle import sampler_oracle as SO
    from oracle import unet_oracle as O
    from stableanimator_b200.arch import UNetConfig
    from stableanimator_b200.pose_net import PoseNetB200
    from stableanimator_b200.sampler import AnimationSampler
    from stableanimator_b200.synth import synth_inputs, synth_posenet_state_dict, synth_unet_state_dict
    from stableanimator_b200.unet import UNetB200
    cfg = UNetConfig.tiny()
    frames, h, w = 6, 16, 16
    sd, psd = synth_unet_state_dict(cfg), synth_posenet_state_dict(cfg.block_out_channels[0])
    sdc, psdc = {k: v.cuda() for k, v in sd.items()}, {k: v.cuda() for k, v in psd.items()}
    inp = {k: v.cuda() for k, v in synth_inputs(cfg, frames, h, w).items()}
    sampler = AnimationSampler(UNetB200(cfg, sd, "cuda", dtype=H), PoseNetB200(psd, cfg.block_out_channels[0], "cuda", dtype=H),
                               num_inference_steps=25, tile_size=4, tile_overlap=2)
    lat = inp["latents"][:, :4].contiguous()
    out = sampler(lat, inp["image_latents"], inp["encoder_hidden_states"], inp["added_time_ids"], inp["pose"])
    ref = SO.denoise_loop(lambda x, t, e, a, pl: O.unet_forward(sdc, cfg, x, t.cuda(), e, a, pl), lambda pm: O.pose_net(psdc, pm),
                          lat, inp["image_latents"], inp["encoder_hidden_states"], inp["added_time_ids"], inp["pose"],
                          num_steps=25, tile_size=4, tile_overlap=2)
    e = rel_l2(out, ref)
    print(f"fp16 mode, tiny UNet, 6 frames in 2 tiles, 25 steps: final-latent rel-L2 {e:.3e}")
    assert torch.isfinite(out).all() and e <= 1e-2
