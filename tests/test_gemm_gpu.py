"""GPU parity of the tapped tcgen05 GEMM (sa_gemm_tap) against fp32 torch ops on the same
bf16-rounded inputs.  Tolerance: see tests/util.py (rel-L2 and excess max error <= 2e-3)."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close, rel_l2

pytestmark = pytest.mark.gpu


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("m,k,n,tile", [
    (128, 64, 64, 64), (256, 128, 128, 128), (1000, 320, 320, 160), (4096, 1280, 1280, 256),
    (384, 192, 64, 32), (777, 640, 640, 0), (65536, 320, 960, 0), (1024, 1280, 1280, 0),
    (130, 1024, 8, 32),
])
def test_plain_linear(m, k, n, tile):
    from stableanimator_b200 import ops
    a, w = rnd(m, k, seed=1), rnd(n, k, scale=k ** -0.5, seed=2)
    bias = torch.randn(n, device="cuda")
    out = ops.gemm_tap(a, w, bias=bias, tile_n=tile)
    ref = a.float() @ w.float().t() + bias
    torch.cuda.synchronize()
    assert out.shape == (m, n)
    assert_close(out, ref)


def test_epilogue_all():
    from stableanimator_b200 import ops
    m, k, n, rpg = 2048, 320, 320, 512
    a, w = rnd(m, k, seed=3), rnd(n, k, scale=k ** -0.5, seed=4)
    bias = torch.randn(n, device="cuda")
    rb = torch.randn(m // rpg, n, device="cuda")
    rs = torch.rand(m // rpg, device="cuda") + 0.5
    res, bx = rnd(m, n, seed=5), rnd(m, n, seed=6)
    alpha = 0.62
    out = ops.gemm_tap(a, w, bias=bias, row_bias=rb, row_scale=rs, rows_per_group=rpg,
                       residual=res, blend_x=bx, alpha=alpha)
    acc = a.float() @ w.float().t()
    g = torch.arange(m, device="cuda") // rpg
    v = acc * rs[g][:, None] + bias + rb[g] + res.float()
    ref = alpha * bx.float() + (1 - alpha) * v
    assert_close(out, ref)


@pytest.mark.parametrize("c,tile", [(64, 128), (320, 0), (320, 256)])
def test_geglu(c, tile):
    from stableanimator_b200 import ops
    m = 1536
    a = rnd(m, c, seed=7)
    w, b = rnd(8 * c, c, scale=c ** -0.5, seed=8), torch.randn(8 * c, device="cuda")
    wi, bi = ops.interleave_geglu(w, b)
    out = ops.gemm_tap(a, wi, bias=bi, geglu=True, tile_n=tile)
    proj = a.float() @ w.float().t() + b
    val, gate = proj.chunk(2, dim=-1)
    ref = val * F.gelu(gate)
    assert out.shape == (m, 4 * c)
    assert_close(out, ref)


def test_strided_a_and_out():
    from stableanimator_b200 import ops
    m, c = 512, 128
    qkv = rnd(m, 3 * c, seed=9)
    w = rnd(c, c, scale=c ** -0.5, seed=10)
    big = torch.zeros(m, 2 * c, device="cuda", dtype=torch.bfloat16)
    ops.gemm_tap(qkv[:, c:2 * c], w, out=big[:, c:])
    ref = qkv[:, c:2 * c].float() @ w.float().t()
    assert_close(big[:, c:], ref)
    assert big[:, :c].abs().max().item() == 0


@pytest.mark.parametrize("f,h,w_,cin,cout", [
    (2, 8, 8, 64, 64), (3, 16, 16, 128, 64), (2, 32, 32, 64, 128), (16, 64, 64, 320, 320),
    (4, 4, 4, 256, 256), (1, 18, 9, 64, 64), (2, 72, 36, 64, 64),
])
def test_conv3x3(f, h, w_, cin, cout):
    from stableanimator_b200 import ops
    x = rnd(f, h, w_, cin, seed=11)                       # NHWC
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=12)  # OIHW like nn.Conv2d
    bias = torch.randn(cout, device="cuda")
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()  # [O, (ky,kx,ci)]
    out = ops.gemm_tap(x, wp, grid=(w_, h, f), taps=ops.TAPS_3X3, bias=bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert_close(out.reshape(f, h, w_, cout), ref)


@pytest.mark.parametrize("f,h,w_,cin,cout,res,stats", [
    (5, 8, 8, 256, 320, True, True),      # 3 row tiles (odd): the pair's second CTA runs a phantom tile in the last item
    (16, 16, 16, 640, 1280, True, False),  # 32 row tiles x 8 column tiles, K = 5760
    (3, 18, 9, 192, 160, False, True),    # ragged boxes at the 9-wide level of the 576x1024 geometry, K = 1728
    (7, 16, 8, 256, 128, True, False),    # tile_n = 128 pair kernel
])
def test_long_k_conv_on_the_cta_pair_kernels(f, h, w_, cin, cout, res, stats):
    """K >= 1280 routes to the cta_group::2 kernels (M = 256 per CTA pair, weight tile split across the pair): every
    epilogue feature they carry, odd row-tile counts and ragged boxes, against F.conv2d."""
    from stableanimator_b200 import ops
    x = rnd(f, h, w_, cin, seed=51)
    wt = rnd(cout, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=52)
    bias = torch.randn(cout, device="cuda")
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    r = rnd(f * h * w_, cout, seed=53) if res else None
    st = torch.empty((f, 32, 2), device="cuda", dtype=torch.float64) if stats else None
    out = ops.gemm_tap(x, wp, grid=(w_, h, f), taps=ops.TAPS_3X3, bias=bias, residual=r, out_scale=0.5,
                       gn_stats=st, gn_rows_per_stat=h * w_ if stats else 0)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, cout) * 0.5
    if res:
        ref = ref + r.float()
    assert_close(out, ref)
    if stats:
        v = out.double().reshape(f, h * w_, 32, cout // 32)
        assert torch.allclose(st[..., 0], v.sum(dim=(1, 3)), rtol=2e-6, atol=1e-4)
        assert torch.allclose(st[..., 1], (v * v).sum(dim=(1, 3)), rtol=2e-6, atol=1e-4)


def test_conv3x3_concat_sources():
    from stableanimator_b200 import ops
    f, h, w_, c0, c1, cout = 2, 16, 16, 128, 64, 128
    x0, x1 = rnd(f, h, w_, c0, seed=13), rnd(f, h, w_, c1, seed=14)
    wt = rnd(cout, c0 + c1, 3, 3, scale=(9 * (c0 + c1)) ** -0.5, seed=15)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    out = ops.gemm_tap(x0, wp, a1=x1, grid=(w_, h, f), taps=ops.TAPS_3X3)
    xin = torch.cat([x0, x1], dim=-1).float().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, wt.float(), None, padding=1).permute(0, 2, 3, 1)
    assert_close(out.reshape(f, h, w_, cout), ref)


@pytest.mark.parametrize("b,f,s,c", [(1, 8, 64, 64), (2, 16, 256, 128), (1, 16, 4096, 320), (1, 4, 16, 64)])
def test_temporal_conv(b, f, s, c):
    from stableanimator_b200 import ops
    x = rnd(b, f, s, c, seed=16)
    wt = rnd(c, c, 3, 1, 1, scale=(3 * c) ** -0.5, seed=17)  # OIDHW like nn.Conv3d
    bias = torch.randn(c, device="cuda")
    wp = wt[:, :, :, 0, 0].permute(0, 2, 1).reshape(c, 3 * c).contiguous()  # [O, (kt, ci)]
    out = ops.gemm_tap(x, wp, grid=(s, f, b), taps=ops.TAPS_T3, bias=bias)
    xin = x.float().permute(0, 3, 1, 2)[..., None]  # [B, C, F, S, 1]
    ref = F.conv3d(xin, wt.float(), bias, padding=(1, 0, 0))[..., 0].permute(0, 2, 3, 1)
    assert_close(out.reshape(b, f, s, c), ref)


def test_fp32_out_small_n():
    from stableanimator_b200 import ops
    f, h, w_, cin = 2, 16, 16, 64
    x = rnd(f, h, w_, cin, seed=18)
    wt = rnd(4, cin, 3, 3, scale=(9 * cin) ** -0.5, seed=19)
    wp = torch.zeros(8, 9 * cin, device="cuda", dtype=torch.bfloat16)
    wp[:4] = wt.permute(0, 2, 3, 1).reshape(4, -1)
    out = ops.gemm_tap(x, wp, grid=(w_, h, f), taps=ops.TAPS_3X3, out_f32=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, padding=1).permute(0, 2, 3, 1)
    assert out.dtype == torch.float32
    assert_close(out.reshape(f, h, w_, 8)[..., :4], ref, tol=1e-4)


def test_out_scale_residual_is_alpha_blend():
    """res-block time mixer: alpha*x + (1-alpha)*(x + conv) == x + (1-alpha)*conv (unet.py:1601-1605)."""
    from stableanimator_b200 import ops
    m, k, n, alpha = 4096, 640, 640, 0.57
    a, w = rnd(m, k, seed=21), rnd(n, k, scale=k ** -0.5, seed=22)
    bias, x = torch.randn(n, device="cuda"), rnd(m, n, seed=23)
    out = ops.gemm_tap(a, w, bias=bias, residual=x, out_scale=1.0 - alpha)
    conv = a.float() @ w.float().t() + bias
    ref = alpha * x.float() + (1 - alpha) * (x.float() + conv)
    assert_close(out, ref)


@pytest.mark.parametrize("m,k,n", [(148 * 128 * 3 + 77, 320, 320), (300, 64, 1280), (128 * 149, 128, 256)])
def test_persistent_many_tiles_and_ragged_rows(m, k, n):
    from stableanimator_b200 import ops
    a, w = rnd(m, k, seed=24), rnd(n, k, scale=k ** -0.5, seed=25)
    res = rnd(m, n, seed=26)
    out = ops.gemm_tap(a, w, residual=res)
    assert_close(out, a.float() @ w.float().t() + res.float())


def test_geglu_matches_exact_erf_gelu_on_wide_range():
    from stableanimator_b200 import ops
    m, c = 512, 64
    a = rnd(m, c, scale=3.0, seed=27)
    w, b = rnd(8 * c, c, scale=1.0, seed=28), torch.randn(8 * c, device="cuda") * 2
    wi, bi = ops.interleave_geglu(w, b)
    out = ops.gemm_tap(a, wi, bias=bi, geglu=True)
    proj = a.float() @ w.float().t() + b
    val, gate = proj.chunk(2, dim=-1)
    assert_close(out, val * F.gelu(gate))


def test_fp16_column_block():
    """f16_from_col: the q|k|v projection stores its V block as IEEE fp16 (saturating) for the fp16 P V
    MMA of sa_attn_spatial; the other columns stay bf16."""
    from stableanimator_b200 import ops
    m, k, c = 1000, 320, 320
    a, w = rnd(m, k, seed=11), rnd(3 * c, k, scale=k ** -0.5, seed=12)
    w[2 * c + 5] *= 1e6  # one V column overflows fp16: must saturate to +-65504, not inf
    out = ops.gemm_tap(a, w, f16_from_col=2 * c)
    ref = a.float() @ w.float().t()
    assert out.dtype == torch.bfloat16
    assert_close(out[:, :2 * c], ref[:, :2 * c], what="q|k (bf16)")
    v = out[:, 2 * c:].contiguous().view(torch.float16).float()
    assert torch.isfinite(v).all()
    vref = ref[:, 2 * c:].clamp(-65504.0, 65504.0)
    # fp16 output: half-ulp is 2^-11 relative; 2^-10 leaves room for the fp32 accumulation-order difference
    assert ((v - vref).abs() <= 2.0 ** -10 * vref.abs() + 1e-5).all()
    with pytest.raises(Exception):
        ops.gemm_tap(a, w, f16_from_col=2 * c + 8)


@pytest.mark.parametrize("f,h,w_,cin,cout,per_frame", [
    (4, 16, 16, 64, 64, True),       # cpg 2, uniform warps
    (4, 16, 16, 128, 320, True),     # cpg 10: groups straddle the 32-column chunks
    (4, 16, 16, 128, 320, False),    # whole-tile statistics (temporal GroupNorm)
    (2, 8, 8, 64, 640, True),        # S = 64: two frames per 128-row tile
    (3, 18, 9, 64, 1280, True),      # cpg 40 > 32; ragged boxes (rows outside the grid, mixed frames per warp)
    (2, 72, 36, 64, 320, True),      # the 576x1024 geometry one level down
])
def test_conv3x3_fused_groupnorm_statistics(f, h, w_, cin, cout, per_frame):
    """sa_gemm_args.gn_stats: Sum / Sum^2 of the bf16 output per (statistic, group) accumulated by the epilogue equal
    the sums over the tensor the launch wrote, and sa_groupnorm_apply on them equals the two-pass sa_groupnorm."""
    from stableanimator_b200 import ops
    x = rnd(f, h, w_, cin, seed=31)
    wp = rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5, seed=32)
    bias = torch.randn(cout, device="cuda")
    res = rnd(f * h * w_, cout, seed=33)
    rps = h * w_ if per_frame else f * h * w_
    ns = f * h * w_ // rps
    st = torch.full((ns, 32, 2), 7.0, device="cuda", dtype=torch.float64)   # the call zeroes it
    plain = ops.gemm_tap(x, wp, grid=(w_, h, f), taps=ops.TAPS_3X3, bias=bias, residual=res, out_scale=0.5)
    out = ops.gemm_tap(x, wp, grid=(w_, h, f), taps=ops.TAPS_3X3, bias=bias, residual=res, out_scale=0.5,
                       gn_stats=st, gn_rows_per_stat=rps)
    assert torch.equal(out, plain)                                          # the output itself is untouched
    v = out.double().reshape(ns, rps, 32, cout // 32)
    ref = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1)
    assert torch.allclose(st, ref, rtol=2e-6, atol=1e-6 * ref.abs().max().item()), (st - ref).abs().max().item()
    gamma, beta = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    two_pass = ops.groupnorm(out, gamma, beta, 1e-5, rps, silu=True)
    fused = ops.groupnorm(out, gamma, beta, 1e-5, rps, silu=True, stats_ws=st, stats_ready=True)
    assert (two_pass.float() - fused.float()).abs().max().item() <= 2.0 ** -7 * two_pass.float().abs().max().item()


def test_fused_groupnorm_statistics_plain_and_rejections():
    """A launch with no other epilogue feature still produces statistics (it leaves the PLAIN instantiation);
    GEGLU / fp32 output / fp16 column blocks are rejected loudly."""
    from stableanimator_b200 import ops
    from stableanimator_b200._lib import SaKernelError
    a, w = rnd(512, 128, seed=41), rnd(256, 128, scale=128 ** -0.5, seed=42)
    st = torch.empty((4, 32, 2), device="cuda", dtype=torch.float64)
    out = ops.gemm_tap(a, w, gn_stats=st, gn_rows_per_stat=128)
    v = out.double().reshape(4, 128, 32, 8)
    assert torch.allclose(st[..., 0], v.sum(dim=(1, 3)), rtol=2e-6, atol=1e-5)
    with pytest.raises(SaKernelError):
        ops.gemm_tap(a, w, geglu=True, gn_stats=torch.empty((4, 32, 2), device="cuda", dtype=torch.float64),
                     gn_rows_per_stat=128)
    with pytest.raises(SaKernelError):
        ops.gemm_tap(a, w, out_f32=True, gn_stats=st, gn_rows_per_stat=128)


@pytest.mark.parametrize("f,h,w_,c,cout", [(2, 8, 8, 64, 64), (3, 16, 12, 128, 64), (2, 9, 18, 64, 128), (16, 32, 32, 640, 640)])
def test_upsample_conv_as_phase_convolutions(f, h, w_, c, cout):
    """Upsample2D (nearest x2 + conv3x3, unet.py:1789-1807) as four 2x2-tap launches writing every other pixel
    (strided output grid, sa_gemm_args.ldc2 / ldc3) vs conv2d(interpolate(x)) and vs the upsample2x + 3x3 path."""
    from stableanimator_b200 import ops
    x = rnd(f, h, w_, c, seed=51)
    wt = rnd(cout, c, 3, 3, scale=(9 * c) ** -0.5, seed=52)
    bias = torch.randn(cout, device="cuda")
    out = torch.full((f, 2 * h, 2 * w_, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    for (py, px, taps), (_, _, taps2, wp) in zip(ops.UPSAMPLE_PHASE_TAPS, ops.upsample_conv_phases(wt)):
        assert taps == taps2
        ops.gemm_tap(x, wp.bfloat16(), grid=(w_, h, f), taps=taps, bias=bias, out=out[:, py::2, px::2, :])
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, wt.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert_close(out, ref)
    wp9 = wt.permute(0, 2, 3, 1).reshape(cout, 9 * c).contiguous()
    old = ops.gemm_tap(ops.upsample2x(x), wp9, grid=(2 * w_, 2 * h, f), taps=ops.TAPS_3X3, bias=bias)
    assert rel_l2(out.float(), old.reshape(out.shape).float()) <= 2.0 ** -8


@pytest.mark.parametrize("f,h,w_,c,cout", [(2, 8, 8, 64, 64), (3, 16, 24, 128, 64), (2, 18, 36, 64, 128), (16, 64, 64, 320, 320)])
def test_stride2_conv_through_tma_element_strides(f, h, w_, c, cout):
    """Downsample2D (Conv2d 3x3, stride 2, pad 1; unet.py:1459-1483) straight from the full-resolution tensor
    (sa_gemm_args.in_stride = 2) vs F.conv2d and vs the im2col_s2 + GEMM path (same operands -> same bits)."""
    from stableanimator_b200 import ops
    x = rnd(f, h, w_, c, seed=61)
    wt = rnd(cout, c, 3, 3, scale=(9 * c) ** -0.5, seed=62)
    bias = torch.randn(cout, device="cuda")
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * c).contiguous()
    out = ops.gemm_tap(x, wp, grid=(w_ // 2, h // 2, f), taps=ops.TAPS_3X3, in_stride=2, bias=bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2, padding=1).permute(0, 2, 3, 1)
    assert_close(out.reshape(f, h // 2, w_ // 2, cout), ref)
    old = ops.gemm_tap(ops.im2col_s2(x), wp, bias=bias)
    assert rel_l2(out.float(), old.float()) <= 1e-3
    with pytest.raises(ValueError):
        ops.gemm_tap(x, wp, taps=ops.TAPS_3X3, in_stride=2, bias=bias)      # grid is mandatory
