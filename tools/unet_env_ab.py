"""A/B of env-selected UNet options in ONE process (graphs replayed alternately): python tools/unet_env_ab.py ATTR
where ATTR is a boolean attribute of UNetB200 (upsample_phases, fuse_gn_stats, skip_zero_context) or env:NAME=v1,v2,...
for an environment switch the library reads per launch (env:SA_PDL=0,1,2)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict  # noqa: E402
from stableanimator_b200.unet import UNetB200  # noqa: E402

attr = sys.argv[1]
h = int(sys.argv[2]) if len(sys.argv) > 2 else 64
w = int(sys.argv[3]) if len(sys.argv) > 3 else h
cfg = UNetConfig.full()
unet = UNetB200(cfg, synth_unet_state_dict(cfg), "cuda")
B, F = 2, 16
inp = synth_inputs(cfg, F, h, w)
ati = inp["added_time_ids"][:B].cuda()
x = torch.randn(B * F, h, w, 8, device="cuda").bfloat16()
t = torch.tensor(1.0, device="cuda")
graphs, outs, ctxs = {}, {}, {}   # ctxs: the graphs read these buffers, keep them alive
vals = (True, False)
if attr.startswith("env:"):
    attr, _, vs = attr[4:].partition("=")
    vals = tuple(vs.split(","))
    env = True
else:
    env = False
for val in vals:
    if env:
        os.environ[attr] = val
    else:
        setattr(unet, attr, val)
    ctx = ctxs[val] = unet.prepare_context(inp["encoder_hidden_states"][:B].cuda())
    unet.forward_nhwc(x, t, ctx, ati, B, F, h, w)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs[val] = unet.forward_nhwc(x, t, ctx, ati, B, F, h, w)
    graphs[val] = g
times = {k: [] for k in graphs}
for _ in range(12):
    for val, g in graphs.items():
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        times[val].append(e0.elapsed_time(e1) / 5)
for val, ts in times.items():
    ts = sorted(ts)
    print(f"{attr}={val!s:5s}: median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f}  max {ts[-1]:.3f}")
ref = outs[vals[-1]][:, :4].double()
for val in vals[:-1]:
    d = (outs[val][:, :4].double() - ref).norm() / ref.norm()
    print(f"output rel-L2 {attr}={val} vs {attr}={vals[-1]}: {d.item():.3e}")
