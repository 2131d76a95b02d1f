"""GPU parity of sa_ff_fused -- LayerNorm -> GEGLU feed-forward -> residual (-> AlphaBlender) in one tcgen05 kernel --
against an fp32 torch restatement of the reference ops (nn.LayerNorm, FeedForward / GEGLU unet.py:827-870, residual add
:553-555 / :936-940 / :949-953, time_mixer :693-697), and against the unfused chain of kernels it replaces."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close, rel_l2

pytestmark = pytest.mark.gpu


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).bfloat16()


def make(c, m, seed, with_addv, groups=1):
    x = rnd(m, c, seed=seed)
    w1, b1 = rnd(8 * c, c, scale=c ** -0.5, seed=seed + 1), torch.randn(8 * c, device="cuda") * 0.5
    w2, b2 = rnd(c, 4 * c, scale=(4 * c) ** -0.5, seed=seed + 2), torch.randn(c, device="cuda") * 0.5
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
    addv = torch.randn(groups, c, device="cuda") if with_addv else None
    return x, w1, b1, w2, b2, gamma, beta, addv


def reference(x, w1, b1, w2, b2, gamma, beta, addv, rows_per_group, blend_x=None, alpha=0.0):
    xr = x.float()
    if addv is not None:
        xr = (xr + addv.repeat_interleave(rows_per_group, dim=0)[: x.shape[0]]).bfloat16().float()
    a = F.layer_norm(xr, (x.shape[1],), gamma, beta, 1e-5).bfloat16().float()   # the unfused path rounds LN to bf16 too
    proj = a @ w1.float().t() + b1
    val, gate = proj.chunk(2, dim=-1)
    h = (val * F.gelu(gate)).bfloat16().float()                                  # hidden tensor is bf16 in both paths
    out = h @ w2.float().t() + b2 + xr
    if blend_x is not None:
        out = alpha * blend_x.float() + (1 - alpha) * out
    return out


@pytest.mark.parametrize("c,m,with_addv,blend", [
    (64, 256, False, False), (128, 1000, True, False), (256, 640, True, True), (320, 4096, False, False),
    (320, 128 * 37 + 5, True, True), (320, 131072, True, True), (64, 77, True, False)])
def test_ff_fused_matches_reference_ops(c, m, with_addv, blend):
    from stableanimator_b200 import ops
    groups = 4 if with_addv else 1
    rpg = -(-m // groups)
    x, w1, b1, w2, b2, gamma, beta, addv = make(c, m, 100 + c, with_addv, groups)
    wi, bi = ops.interleave_geglu(w1, b1)
    bx = rnd(m, c, seed=5) if blend else None
    out = ops.ff_fused(x, gamma, beta, wi, bi, w2, b2, addv=addv, rows_per_group=rpg, blend_x=bx, alpha=0.37)
    ref = reference(x, w1, b1, w2, b2, gamma, beta, addv, rpg, bx, 0.37)
    assert out.shape == (m, c) and out.dtype == torch.bfloat16
    assert_close(out, ref, what=f"ff_fused C={c} M={m}")


def test_ff_fused_equals_the_unfused_kernel_chain():
    """Same result (to bf16 rounding of the same intermediates) as sa_layernorm -> sa_gemm_tap(geglu) ->
    sa_gemm_tap(residual, blend) -- the launches it replaces in UNetB200._tst."""
    from stableanimator_b200 import ops
    c, m = 320, 16 * 1024
    x, w1, b1, w2, b2, gamma, beta, addv = make(c, m, 7, True, 2)
    wi, bi = ops.interleave_geglu(w1, b1)
    bx = rnd(m, c, seed=9)
    fused = ops.ff_fused(x, gamma, beta, wi, bi, w2, b2, addv=addv, rows_per_group=m // 2, blend_x=bx, alpha=0.25)
    z0 = torch.empty_like(x)
    a = ops.layernorm(x, gamma, beta, addv=addv, rows_per_group=m // 2, x_out=z0)
    g = ops.gemm_tap(a, wi, bias=bi, geglu=True)
    chain = ops.gemm_tap(g, w2, bias=b2, residual=z0, blend_x=bx, alpha=0.25)
    assert rel_l2(fused, chain) <= 2e-3
    assert (fused.float() - chain.float()).abs().max().item() <= 2.0 ** -6 * chain.float().abs().max().item()


def test_ff_fused_strided_views_and_rejections():
    from stableanimator_b200 import ops
    from stableanimator_b200._lib import SaKernelError
    c, m = 128, 512
    x, w1, b1, w2, b2, gamma, beta, _ = make(c, m, 3, False)
    wi, bi = ops.interleave_geglu(w1, b1)
    big_in = torch.zeros(m, 2 * c, device="cuda", dtype=torch.bfloat16)
    big_in[:, c:] = x
    big_out = torch.zeros(m, 3 * c, device="cuda", dtype=torch.bfloat16)
    ops.ff_fused(big_in[:, c:], gamma, beta, wi, bi, w2, b2, out=big_out[:, c:2 * c])
    assert_close(big_out[:, c:2 * c], reference(x, w1, b1, w2, b2, gamma, beta, None, 1))
    assert big_out[:, :c].abs().max().item() == 0 and big_out[:, 2 * c:].abs().max().item() == 0
    with pytest.raises(ValueError):
        ops.ff_fused(rnd(64, 192), gamma, beta, wi, bi, w2, b2)
    with pytest.raises((ValueError, SaKernelError)):
        ops.ff_fused(x, gamma, beta, wi[:, :64].contiguous(), bi, w2, b2)
