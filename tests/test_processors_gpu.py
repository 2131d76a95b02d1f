"""GPU parity of the drop-in attention processors (stableanimator_b200/processors.py) called exactly
as the reference's ``Attention.forward`` calls them (unet.py:764-774), against the oracle's processor
restatements.  ``Attention`` below mirrors the attributes of the reference module (unet.py:720-757)."""
import pytest
import torch
import torch.nn as nn

from util import assert_close

# a processor call chains three kernels (projection GEMM -> attention -> to_out GEMM) with bf16
# intermediates; each is within the per-kernel 2e-3 (tests/test_gemm_gpu.py, test_kernels_gpu.py), the
# chain is bounded by twice that
CHAIN_TOL = 4e-3

pytestmark = pytest.mark.gpu


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=5, dim_head=64, processor=None):
        super().__init__()
        self.head_size = self.heads = heads
        self.hidden_size = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, self.hidden_size, bias=False)
        self.to_k = nn.Linear(self.cross_attention_dim, self.hidden_size, bias=False)
        self.to_v = nn.Linear(self.cross_attention_dim, self.hidden_size, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(self.hidden_size, query_dim, bias=True), nn.Dropout(0.0)])
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states)


def make(query_dim, ctx_dim, heads, proc, seed):
    torch.manual_seed(seed)
    a = Attention(query_dim, ctx_dim, heads, processor=proc).cuda()
    with torch.no_grad():
        for p in list(a.parameters()) + (list(proc.parameters()) if isinstance(proc, nn.Module) else []):
            p.copy_(p.to(torch.bfloat16).float())  # bf16-representable weights
    if isinstance(proc, nn.Module):
        proc.cuda()
    return a


def state(a, proc=None):
    sd = {f"a.{k}": v.detach().float() for k, v in a.state_dict().items() if not k.startswith("processor")}
    if proc is not None and isinstance(proc, nn.Module):
        sd.update({f"a.processor.{k}": v.detach().float() for k, v in proc.state_dict().items()})
    return sd


@pytest.mark.parametrize("frames,s,heads,dtype", [(4, 256, 2, torch.bfloat16), (2, 1024, 5, torch.float32),
                                                   (3, 64, 4, torch.float16)])
def test_animation_attn_processor(frames, s, heads, dtype):
    from oracle import unet_oracle as O
    from stableanimator_b200.processors import AnimationAttnProcessor
    c = heads * 64
    a = make(c, None, heads, AnimationAttnProcessor(hidden_size=c), 1)
    x = torch.randn(frames, s, c, device="cuda").to(torch.bfloat16)
    out = a(x.to(dtype))
    assert out.dtype == dtype and out.shape == x.shape
    ref = O.attn_self(state(a), "a", x.float(), heads)
    assert_close(out.to(torch.bfloat16), ref, tol=CHAIN_TOL, what="AnimationAttnProcessor")


@pytest.mark.parametrize("frames,s,heads", [(4, 256, 2), (2, 1024, 10)])
def test_animation_id_attn_processor(frames, s, heads):
    from oracle import unet_oracle as O
    from stableanimator_b200.processors import AnimationIDAttnNormalizedProcessor
    c = heads * 64
    proc = AnimationIDAttnNormalizedProcessor(hidden_size=c, cross_attention_dim=1024)
    a = make(c, 1024, heads, proc, 2)
    assert sorted(k for k, _ in proc.named_parameters()) == ["id_to_k.weight", "id_to_v.weight"]
    x = torch.randn(frames, s, c, device="cuda").to(torch.bfloat16)
    ehs = (torch.randn(1, 5, 1024, device="cuda") * 2).to(torch.bfloat16).float().repeat_interleave(frames, 0)
    out = a(x, encoder_hidden_states=ehs)
    ref = O.attn_id_normalized(state(a, proc), "a", x.float(), ehs, heads)
    assert_close(out, ref, tol=CHAIN_TOL, what="AnimationIDAttnNormalizedProcessor")
    with pytest.raises(AssertionError):
        AnimationIDAttnNormalizedProcessor(hidden_size=c, cross_attention_dim=768)


@pytest.mark.parametrize("bs,f,heads", [(256, 16, 2), (1024, 8, 5)])
def test_xformers_attn_processor_self_and_cross(bs, f, heads):
    from oracle import unet_oracle as O
    from stableanimator_b200.processors import XFormersAttnProcessor
    c = heads * 64
    a = make(c, None, heads, XFormersAttnProcessor(), 3)
    x = torch.randn(bs, f, c, device="cuda").to(torch.bfloat16)
    out = a(x)
    assert_close(out, O.attn_self(state(a), "a", x.float(), heads), tol=CHAIN_TOL, what="XFormers self")
    a2 = make(c, 1024, heads, XFormersAttnProcessor(), 4)
    ctx = torch.randn(1, 1, 1024, device="cuda").to(torch.bfloat16).expand(bs, 1, 1024).contiguous()
    out2 = a2(x, encoder_hidden_states=ctx)
    assert_close(out2, O.attn_cross(state(a2), "a", x.float(), ctx.float(), heads), tol=CHAIN_TOL, what="XFormers cross")
    with pytest.raises(AssertionError):
        XFormersAttnProcessor(attention_op="cutlass")
