"""Host-side algebra that the CUDA path relies on, checked on CPU against plain torch (no kernels involved)."""
import pytest
import torch
import torch.nn.functional as F


@pytest.mark.parametrize("h,w,c,o", [(5, 6, 5, 3), (1, 1, 2, 2), (8, 3, 4, 6)])
def test_upsample_conv_phase_decomposition(h, w, c, o):
    """ops.upsample_conv_phases: nearest x2 + conv3x3 (unet.py:1789-1807) == four 2x2-tap convolutions of the
    low-resolution input with summed taps, zero padding included."""
    from stableanimator_b200.ops import UPSAMPLE_PHASE_TAPS, upsample_conv_phases
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, c, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(o, c, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(o, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), wt, b, padding=1)
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for (py, px, taps), (py2, px2, taps2, wp) in zip(UPSAMPLE_PHASE_TAPS, upsample_conv_phases(wt)):
        assert (py, px, taps) == (py2, px2, taps2) and wp.shape == (o, 4 * c)
        acc = torch.zeros(2, o, h, w, dtype=torch.float64)
        for t, (dx, dy, dz) in enumerate(taps):
            assert dz == 0
            acc += torch.einsum("oi,bihw->bohw", wp[:, t * c:(t + 1) * c].double(),
                                xp[:, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w])
        out[:, :, py::2, px::2] = acc + b.view(1, o, 1, 1)
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()   # weights are summed in fp32


def test_geglu_interleave_roundtrip():
    """ops.interleave_geglu: 64-row value / gate blocks; de-interleaving restores [value ; gate] (unet.py:865-870)."""
    from stableanimator_b200.ops import interleave_geglu
    w = torch.arange(256 * 3, dtype=torch.float32).reshape(256, 3)
    b = torch.arange(256, dtype=torch.float32)
    wi, bi = interleave_geglu(w, b)
    blocks = wi.reshape(2, 2, 64, 3)            # [pair, value|gate, 64, k]
    assert torch.equal(blocks[:, 0].reshape(128, 3), w[:128]) and torch.equal(blocks[:, 1].reshape(128, 3), w[128:])
    assert torch.equal(bi.reshape(2, 2, 64)[:, 0].reshape(-1), b[:128])


def test_zero_context_attention_is_the_output_bias():
    """The shortcut UNetB200._tst takes for the unconditional CFG branch: with an all-zero context the reference's
    AnimationIDAttnNormalizedProcessor (unet.py:1938-2029; oracle restatement) returns exactly to_out's bias."""
    from oracle import unet_oracle as O
    from stableanimator_b200.arch import UNetConfig, build_stages
    from stableanimator_b200.synth import synth_unet_state_dict
    cfg = UNetConfig.tiny()
    sd = synth_unet_state_dict(cfg)
    t = next(t for st in build_stages(cfg) for t in st.tst if t is not None)
    name = f"{t.name}.transformer_blocks.0.attn2"
    x = torch.randn(3, 16, t.c, generator=torch.Generator().manual_seed(1))
    out = O.attn_id_normalized(sd, name, x, torch.zeros(3, 1 + cfg.num_tokens, cfg.cross_attention_dim), t.heads)
    bias = sd[f"{name}.to_out.0.bias"]
    assert torch.allclose(out, bias.expand_as(out), atol=1e-6)
