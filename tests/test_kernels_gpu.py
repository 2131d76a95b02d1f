"""GPU parity of the non-GEMM kernels against the oracle's fp32 functions (oracle/unet_oracle.py) or
plain fp32 torch ops on the same bf16-rounded inputs.  Tolerances: tests/util.py."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import assert_close

pytestmark = pytest.mark.gpu


def rnd(*shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


# ------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("frames,s,heads", [(2, 256, 2), (1, 64, 4), (3, 1024, 1), (2, 16, 2), (1, 576, 2),
                                            (16, 4096, 5)])
def test_attn_spatial(frames, s, heads):
    from oracle import unet_oracle as O
    from stableanimator_b200 import ops
    c = heads * 64
    qkv = rnd(frames * s, 3 * c, seed=1)
    q, k, v = (O.heads_split(t.float().reshape(frames, s, c), heads) for t in qkv.chunk(3, dim=-1))
    ops.v_block_to_f16_(qkv, c)  # the kernel takes v as fp16 bit patterns (exact re-encoding of these values)
    out = ops.attn_spatial(qkv, frames, s, heads)
    ref = O.heads_merge(O.sdpa(q, k, v), heads).reshape(frames * s, c)
    assert_close(out, ref, what="attn_spatial")


@pytest.mark.parametrize("b,f,s,heads", [(1, 16, 64, 2), (2, 8, 100, 1), (1, 4, 16, 4), (1, 16, 4096, 5)])
def test_attn_temporal(b, f, s, heads):
    from oracle import unet_oracle as O
    from stableanimator_b200 import ops
    c = heads * 64
    qkv = rnd(b * f * s, 3 * c, seed=2)
    out = ops.attn_temporal(qkv, b, f, s, heads)
    # reference permute [B*F,S,C] -> [B*S,F,C] (unet.py:931-933)
    x = qkv.float().reshape(b, f, s, 3 * c).permute(0, 2, 1, 3).reshape(b * s, f, 3 * c)
    q, k, v = (O.heads_split(t, heads) for t in x.chunk(3, dim=-1))
    o = O.heads_merge(O.sdpa(q, k, v), heads)
    ref = o.reshape(b, s, f, c).permute(0, 2, 1, 3).reshape(b * f * s, c)
    assert_close(out, ref, what="attn_temporal")


@pytest.mark.parametrize("b,f,s,heads", [(1, 4, 256, 2), (2, 3, 64, 1), (1, 2, 16, 4), (2, 16, 1024, 10)])
def test_attn_id_processor(b, f, s, heads):
    """Whole AnimationIDAttnNormalizedProcessor (unet.py:1938-2029) through the product kernels:
    to_q GEMM -> sa_attn_id (+stats) -> finalize -> to_out GEMM with row scale/bias."""
    from oracle import unet_oracle as O
    from stableanimator_b200 import ops
    c = heads * 64
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    x = rnd(b * f, s, c, seed=4)
    ehs_b = (torch.randn(b, 5, 1024, generator=g, device=dev) * 1.5).bfloat16().float()
    sd = {}
    for nm, (o_, i_) in {"to_q": (c, c), "to_k": (c, 1024), "to_v": (c, 1024), "to_out.0": (c, c),
                         "processor.id_to_k": (c, 1024), "processor.id_to_v": (c, 1024)}.items():
        sd[f"a.{nm}.weight"] = (torch.randn(o_, i_, generator=g, device=dev) * i_ ** -0.5).bfloat16().float()
    sd["a.to_out.0.bias"] = (torch.randn(c, generator=g, device=dev) * 0.1).bfloat16().float()
    ehs = ehs_b.repeat_interleave(f, dim=0)
    ref = O.attn_id_normalized(sd, "a", x.float(), ehs, heads).reshape(b * f * s, c)

    w = {k: v.bfloat16() for k, v in sd.items() if k.endswith("weight")}
    xq = ops.gemm_tap(x.reshape(-1, c), w["a.to_q.weight"])
    clip = ehs_b[:, 0].contiguous()                      # [b, 1024]
    face = ehs_b[:, 1:].reshape(b * 4, 1024).contiguous()
    vclip = ops.small_linear(clip, w["a.to_v.weight"])
    ik = ops.small_linear(face, w["a.processor.id_to_k.weight"]).reshape(b, 4, c)
    iv = ops.small_linear(face, w["a.processor.id_to_v.weight"]).reshape(b, 4, c)
    wv = ops.small_linear(vclip, w["a.to_out.0.weight"])
    rowsum = sd["a.to_out.0.weight"].sum(dim=1).contiguous()
    stats = torch.empty(b * f, 2, device=dev, dtype=torch.float64)
    ip = ops.attn_id(xq, ik, iv, b, f, s, heads, stats)
    rs = torch.empty(b * f, device=dev)
    rb = torch.empty(b * f, c, device=dev)
    ops.attn_id_finalize(stats, vclip, wv, rowsum, sd["a.to_out.0.bias"], b, f, s, c, rs, rb)
    out = ops.gemm_tap(ip, w["a.to_out.0.weight"], row_scale=rs, row_bias=rb, rows_per_group=s)
    assert_close(out, ref, what="attn_id_processor")


@pytest.mark.parametrize("b,f,s,heads,view", [(1, 2, 100, 3, "dense"), (2, 16, 4096, 5, "dense"), (1, 3, 37, 1, "offset"),
                                              (2, 2, 520, 20, "strided")])
def test_attn_id_kernel_shapes(b, f, s, heads, view):
    """sa_attn_id alone against softmax(q k^T / 8) v in fp32 torch (unet.py:1999-2010) and its fp64 (Sum, Sum^2):
    ragged S, a thread block that is not 320 threads (C = 192), the many-batches-per-CTA path (131072 rows), the
    one-row-per-thread path, K / V views that are only 4-byte aligned (scalar loads) or strided (ctx.all slices)."""
    from stableanimator_b200 import ops
    c = heads * 64
    q = rnd(b * f * s, c, seed=11)
    g = torch.Generator(device="cuda").manual_seed(12)
    if view == "dense":
        ik, iv = (torch.randn(b, 4, c, generator=g, device="cuda") for _ in range(2))
    elif view == "offset":   # 4-byte-aligned base: the scalar K / V load path
        base = torch.randn(2, b, 4, c + 1, generator=g, device="cuda")
        ik, iv = base[0, :, :, 1:], base[1, :, :, 1:]
    else:                    # slices of one [b, 5, n] context tensor, as UNetB200.prepare_context hands them over
        base = torch.randn(b, 5, 2 * c + 64, generator=g, device="cuda")
        ik, iv = base[:, 1:, 64:64 + c], base[:, 1:, 64 + c:64 + 2 * c]
    stats = torch.empty(b * f, 2, device="cuda", dtype=torch.float64)
    out = ops.attn_id(q, ik, iv, b, f, s, heads, stats)
    qh = q.float().reshape(b, f * s, heads, 64)
    kh, vh = ik.float().reshape(b, 4, heads, 64), iv.float().reshape(b, 4, heads, 64)
    p = torch.softmax(torch.einsum("bqhd,bkhd->bhqk", qh, kh) * 0.125, dim=-1)
    ref = torch.einsum("bhqk,bkhd->bqhd", p, vh).reshape(b * f * s, c)
    assert_close(out, ref, what=f"attn_id {view}")
    r = ref.double().reshape(b * f, -1)
    got = stats.cpu()
    want = torch.stack([r.sum(1), (r * r).sum(1)], 1).cpu()
    assert torch.allclose(got, want, rtol=2e-5, atol=1e-3), (got, want)


def test_attn_id_uncond_zero_context():
    """ehs == 0 (uncond branch): both branches vanish, output == to_out bias."""
    from stableanimator_b200 import ops
    b, f, s, heads = 1, 2, 64, 2
    c = heads * 64
    xq = rnd(b * f * s, c, seed=5)
    z = torch.zeros(b, c, device="cuda")
    zk = torch.zeros(b, 4, c, device="cuda")
    stats = torch.empty(b * f, 2, device="cuda", dtype=torch.float64)
    ip = ops.attn_id(xq, zk, zk, b, f, s, heads, stats)
    rs, rb = torch.empty(b * f, device="cuda"), torch.empty(b * f, c, device="cuda")
    bias = torch.randn(c, device="cuda")
    ops.attn_id_finalize(stats, z, z, torch.randn(c, device="cuda"), bias, b, f, s, c, rs, rb)
    assert ip.abs().max().item() == 0 and rs.abs().max().item() == 0
    assert torch.equal(rb, bias.expand(b * f, c))


# ------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("ns,rows,c,silu", [(4, 256, 64, True), (16, 4096, 320, True), (2, 1024, 1280, False),
                                            (3, 64, 960, True), (1, 16 * 64, 256, True)])
def test_groupnorm(ns, rows, c, silu):
    from stableanimator_b200 import ops
    x = rnd(ns * rows, c, scale=2.0, seed=6) + 0.5
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    out = ops.groupnorm(x, gamma, beta, 1e-6, rows, silu=silu)
    xr = x.float().reshape(ns, rows, c).permute(0, 2, 1)  # [ns, C, rows]
    ref = F.group_norm(xr, 32, gamma, beta, 1e-6)
    if silu:
        ref = F.silu(ref)
    assert_close(out, ref.permute(0, 2, 1).reshape(ns * rows, c), what="groupnorm")


@pytest.mark.parametrize("ns,rows,c0,c1,silu", [
    (8, 4096, 320, 0, True),      # streamed path, per-frame statistics (L0 shape)
    (2, 16 * 1024, 640, 0, True),  # per-tile statistics (TemporalResnetBlock, 5-D GroupNorm)
    (4, 1024, 1280, 1280, True),  # concat, C = 2560: two column sets per thread, 4-row tiles
    (6, 1024, 1280, 640, False),  # concat with a group straddling the two sources (cpg = 60)
    (8, 2304, 640, 320, True),    # 576x1024 level-1 concat (S = 64 x 36)
])
def test_groupnorm_streamed(ns, rows, c0, c1, silu):
    """Shapes large enough (>= 8 MB) to take the bulk-copy streamed kernels (gn_stream_kernel)."""
    from stableanimator_b200 import ops
    x0 = rnd(ns * rows, c0, scale=2.0, seed=16) + 0.5
    x1 = rnd(ns * rows, c1, scale=3.0, seed=17) - 0.25 if c1 else None
    c = c0 + c1
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    out = ops.groupnorm(x0, gamma, beta, 1e-5, rows, x1=x1, silu=silu)
    xc = x0 if x1 is None else torch.cat([x0, x1], dim=-1)
    ref = F.group_norm(xc.float().reshape(ns, rows, c).permute(0, 2, 1), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    assert_close(out, ref.permute(0, 2, 1).reshape(ns * rows, c), what="groupnorm streamed")


@pytest.mark.parametrize("rows,c,addv", [(65536, 320, False), (20000, 640, True), (8192, 1280, True)])
def test_layernorm_streamed(rows, c, addv):
    """>= 8 MB: bulk-copy streamed LayerNorm, with and without the fused vector add / residual output."""
    from stableanimator_b200 import ops
    x = rnd(rows, c, scale=2.0, seed=18) + 0.3
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    if addv:
        rpg = rows // 4
        av = torch.randn(4, c, device="cuda")
        xo = torch.empty_like(x)
        out = ops.layernorm(x, gamma, beta, addv=av, rows_per_group=rpg, x_out=xo)
        xs = (x.float() + av.repeat_interleave(rpg, dim=0))
        assert_close(xo, xs, what="layernorm x_out")
        ref = F.layer_norm(xs.bfloat16().float(), (c,), gamma, beta, 1e-5)
    else:
        out = ops.layernorm(x, gamma, beta)
        ref = F.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
    assert_close(out, ref, what="layernorm streamed")


def test_groupnorm_concat_straddling_groups():
    from stableanimator_b200 import ops
    ns, rows, c0, c1 = 2, 256, 1280, 640  # cpg = 60: group 21 straddles the two sources
    x0, x1 = rnd(ns * rows, c0, seed=7), rnd(ns * rows, c1, scale=3.0, seed=8)
    c = c0 + c1
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    out = ops.groupnorm(x0, gamma, beta, 1e-5, rows, x1=x1, silu=True)
    xr = torch.cat([x0, x1], dim=-1).float().reshape(ns, rows, c).permute(0, 2, 1)
    ref = F.silu(F.group_norm(xr, 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(ns * rows, c)
    assert_close(out, ref, what="groupnorm concat")


@pytest.mark.parametrize("rows,c", [(1000, 64), (4096, 320), (777, 640), (512, 1280), (64, 256)])
def test_layernorm(rows, c):
    from stableanimator_b200 import ops
    x = rnd(rows, c, scale=3.0, seed=9) + 1.0
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    out = ops.layernorm(x, gamma, beta)
    assert_close(out, F.layer_norm(x.float(), (c,), gamma, beta, 1e-5), what="layernorm")


def test_layernorm_add_vector():
    from stableanimator_b200 import ops
    groups, rpg, c = 4, 96, 320
    x = rnd(groups * rpg, c, seed=10)
    addv = torch.randn(groups, c, device="cuda")
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    xo = torch.empty_like(x)
    out = ops.layernorm(x, gamma, beta, addv=addv, rows_per_group=rpg, x_out=xo)
    s = x.float() + addv.repeat_interleave(rpg, dim=0)
    assert_close(xo, s, what="ln x_out")
    assert_close(out, F.layer_norm(xo.float(), (c,), gamma, beta, 1e-5), what="ln add")


# ------------------------------------------------------------------------------ small ops
@pytest.mark.parametrize("m,k,n,si,so", [(1, 320, 1280, False, True), (2, 1280, 1280, True, False),
                                         (10, 1024, 640, False, False), (16, 320, 1280, False, True)])
def test_small_linear(m, k, n, si, so):
    from stableanimator_b200 import ops
    x = torch.randn(m, k, device="cuda")
    w = rnd(n, k, scale=k ** -0.5, seed=11)
    b = torch.randn(n, device="cuda")
    out = ops.small_linear(x, w, b, silu_in=si, silu_out=so)
    xi = F.silu(x) if si else x
    ref = xi @ w.float().t() + b
    if so:
        ref = F.silu(ref)
    assert_close(out, ref, tol=2e-5, what="small_linear")


def test_sinusoid_matches_oracle():
    from oracle import unet_oracle as O
    from stableanimator_b200 import ops
    t = torch.tensor([1.637770, -1.553652, 6.0, 127.0, 0.02, 0.0, 15.0], device="cuda")
    for dim in (320, 256, 64):
        assert_close(ops.sinusoid(t, dim), O.timestep_embedding(t, dim), tol=2e-5, what="sinusoid")


@pytest.mark.parametrize("cin,cout,k,stride,pad", [(8, 320, 3, 1, 1), (3, 3, 3, 1, 1), (3, 16, 4, 2, 1),
                                                   (32, 64, 4, 2, 1), (64, 128, 3, 1, 1)])
def test_conv_direct(cin, cout, k, stride, pad):
    from stableanimator_b200 import ops
    n, h, w = 2, 32, 24
    x = rnd(n, h, w, cin, seed=12)
    wt = rnd(cout, cin, k, k, scale=(cin * k * k) ** -0.5, seed=13)
    b = torch.randn(cout, device="cuda")
    cpad = (cout + 7) // 8 * 8
    wp = torch.zeros(k * k * cin, cpad, device="cuda", dtype=torch.bfloat16)
    wp[:, :cout] = wt.permute(2, 3, 1, 0).reshape(k * k * cin, cout)
    out = ops.conv_direct(x, wp, b, cout, k, stride, pad, silu=True, scale=2.0)
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=stride, padding=pad)) * 2.0
    assert_close(out, ref.permute(0, 2, 3, 1), what="conv_direct")


def test_conv_in_with_pose_add():
    from stableanimator_b200 import ops
    n, h, w = 4, 16, 16
    x = rnd(n, h, w, 8, seed=14)
    wt = rnd(64, 8, 3, 3, scale=72 ** -0.5, seed=15)
    pose = rnd(2, h, w, 64, seed=16)
    wp = wt.permute(2, 3, 1, 0).reshape(72, 64).contiguous()
    out = ops.conv_direct(x, wp, None, 64, 3, 1, 1, add=pose, add_from=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, padding=1).permute(0, 2, 3, 1).clone()
    ref[2:] += pose.float()
    assert_close(out, ref, what="conv_in+pose")


def test_upsample_and_im2col():
    from stableanimator_b200 import ops
    x = rnd(3, 8, 12, 64, seed=17)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    # stride-2 conv through im2col + GEMM vs F.conv2d (Downsample2D, unet.py:1474)
    wt = rnd(128, 64, 3, 3, scale=(9 * 64) ** -0.5, seed=18)
    b = torch.randn(128, device="cuda")
    cols = ops.im2col_s2(x)
    out = ops.gemm_tap(cols, wt.permute(0, 2, 3, 1).reshape(128, -1).contiguous(), bias=b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert_close(out.reshape(3, 4, 6, 128), ref, what="downsample conv")


def test_layout_roundtrip():
    from stableanimator_b200 import ops
    x = torch.randn(3, 8, 5, 7, device="cuda")
    y = ops.nchw_to_nhwc(x)
    assert torch.equal(y.float(), x.bfloat16().float().permute(0, 2, 3, 1))
    z = torch.randn(3 * 35, 8, device="cuda")
    back = ops.nhwc_to_nchw_f32(z, 3, 4, 35)
    assert torch.equal(back, z.reshape(3, 35, 8)[..., :4].permute(0, 2, 1))


def test_sampler_kernels_match_oracle():
    from oracle import sampler_oracle as SO
    from stableanimator_b200 import ops
    F_, T, h, w = 7, 4, 8, 6
    hw = h * w
    lat = torch.randn(F_, 4, h, w, device="cuda") * 100
    img = torch.randn(2, F_, 4, h, w, device="cuda")
    idx = torch.tensor([0, 3, 4, 5], device="cuda", dtype=torch.int32)
    sigma, sigma_next = 12.5, 7.25
    out = torch.empty(2 * T * hw, 8, device="cuda", dtype=torch.bfloat16)
    ops.sampler_prepare(lat, img, idx, 2, 0, F_, hw, sigma, out)
    x = SO.scale_model_input(lat[None].expand(2, -1, -1, -1, -1), sigma)
    ref = torch.cat([x, img], dim=2)[:, idx.long()].permute(0, 1, 3, 4, 2).reshape(-1, 8)
    assert_close(out, ref, what="sampler_prepare")
    # accumulate + step
    uo = torch.randn(2 * T * hw, 8, device="cuda")
    wgt = SO.tile_weights(T).cuda()
    noise = torch.zeros(2, F_, 4, h, w, device="cuda")
    ops.sampler_accumulate(uo, idx, wgt, 2, 0, F_, hw, noise)
    ref_n = torch.zeros_like(noise)
    ref_n[:, idx.long()] += uo.reshape(2, T, h, w, 8)[..., :4].permute(0, 1, 4, 2, 3) * wgt[None, :, None, None, None]
    assert_close(noise, ref_n, tol=1e-6, what="sampler_accumulate")
    cnt = torch.rand(F_, device="cuda") + 0.5
    g = torch.linspace(1.0, 3.0, F_, device="cuda")
    noise = torch.randn(2, F_, 4, h, w, device="cuda")
    lat2 = lat.clone()
    ops.sampler_step(noise, cnt, g, F_, 4 * hw, sigma, sigma_next, lat2)
    nn_ = noise / cnt[None, :, None, None, None]
    eps = nn_[0] + g[:, None, None, None] * (nn_[1] - nn_[0])
    ref_l = SO.euler_step_vpred(eps, lat, sigma, sigma_next)
    assert_close(lat2, ref_l, tol=1e-5, what="sampler_step")
