"""Operator-level drop-in boundary against the REAL reference modules, on CPU.

The unmodified reference UNet (``animation/modules/unet.py``, imported through ``oracle/ref_loader.py`` and the
stubs) must accept OUR processor classes in its 64 ``Attention.processor`` slots through its own
``set_attn_processor`` (unet.py:205-225, :756-757), expose the same ``state_dict`` keys as with its own processors
(including ``...attn2.processor.id_to_{k,v}.weight``), and that ``state_dict`` must strict-load into ``UNetB200``.
Needs the reference checkout (build container only: /root/reference does not exist on the GPU box)."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


def _wire(unet, cfg, mod):
    """inference_basic.py:348-377 with the processor classes taken from ``mod``."""
    boc = cfg.block_out_channels
    procs = {}
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            hidden = boc[-1]
        elif name.startswith("up_blocks"):
            hidden = list(reversed(boc))[int(name[len("up_blocks.")])]
        else:
            hidden = boc[int(name[len("down_blocks.")])]
        if "temporal_transformer_blocks" in name:
            procs[name] = mod.XFormersAttnProcessor()
        elif name.endswith("attn1.processor"):
            procs[name] = mod.AnimationAttnProcessor(hidden_size=hidden)
        else:
            procs[name] = mod.AnimationIDAttnNormalizedProcessor(hidden_size=hidden,
                                                                cross_attention_dim=cfg.cross_attention_dim)
    return procs


@pytest.fixture(scope="module")
def tiny():
    from stableanimator_b200.arch import UNetConfig
    cfg = UNetConfig.tiny()
    torch.manual_seed(0)
    ref = ref_loader.build_reference_unet(cfg)          # the reference with its OWN processors
    return cfg, ref


def test_reference_unet_takes_our_processors(tiny):
    import stableanimator_b200.processors as ours
    cfg, ref = tiny
    ref_unet, _, _ = ref_loader.import_reference()
    torch.manual_seed(0)
    other = ref_loader.build_reference_unet(cfg)
    names = list(other.attn_processors.keys())
    assert len(names) == 64                                             # SURVEY 8b registry
    procs = _wire(other, cfg, ours)
    other.set_attn_processor(procs)                                     # the reference's own setter
    # every slot now holds one of our objects, by the reference's wiring rule
    n_self = n_id = n_temporal = 0
    for name, module in other.named_modules():
        if isinstance(module, ref_unet.Attention):
            p = module.processor
            assert type(p).__module__ == "stableanimator_b200.processors", name
            n_temporal += isinstance(p, ours.XFormersAttnProcessor)
            n_self += isinstance(p, ours.AnimationAttnProcessor)
            n_id += isinstance(p, ours.AnimationIDAttnNormalizedProcessor)
    assert (n_self, n_id, n_temporal) == (16, 16, 32)
    # identical parameter names and shapes, incl. the id_to_k / id_to_v the ID processor owns
    a, b = ref.state_dict(), other.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    assert sum(k.endswith("attn2.processor.id_to_k.weight") for k in b) == 16
    # so a reference checkpoint loads strictly into the reference-with-our-processors ...
    other.load_state_dict(a, strict=True)
    # ... and the count check of set_attn_processor (unet.py:208-212) still fires with our classes
    with pytest.raises(ValueError):
        other.set_attn_processor({k: v for k, v in list(procs.items())[:3]})


def test_reference_state_dict_strict_loads_into_unet_b200(tiny):
    from stableanimator_b200.arch import unet_param_shapes
    from stableanimator_b200.unet import UNetB200
    cfg, ref = tiny
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    shapes = unet_param_shapes(cfg)
    assert set(shapes) == set(sd) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    net = UNetB200(cfg, sd, device="cpu")                               # packing only: no kernel runs on CPU
    assert net.n_temb > 0 and net.n_ctx > 0
    bad = dict(sd)
    bad["bogus.weight"] = torch.zeros(1)
    with pytest.raises(KeyError):
        UNetB200(cfg, bad, device="cpu")
    del sd["conv_in.weight"]
    with pytest.raises(KeyError):
        UNetB200(cfg, sd, device="cpu")


def test_our_processors_refuse_cpu_tensors(tiny):
    """No CPU fallback behind the boundary: the reference forward with our processors installed raises on CPU."""
    import stableanimator_b200.processors as ours
    cfg, _ = tiny
    ref_unet, _, _ = ref_loader.import_reference()
    attn = ref_unet.Attention(query_dim=128, heads=2, dim_head=64)
    attn.set_processor(ours.AnimationAttnProcessor(hidden_size=128))
    with pytest.raises((RuntimeError, ValueError)):
        attn(torch.zeros(1, 16, 128))
