"""Time the B=2 UNet call (CUDA graph) with a given build of the library:
  python tools/unet_lib_ab.py [path/to/libsa_b200.so]     (A/B of two builds: run it alternately)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stableanimator_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    import ctypes
    _lib.LIB_PATH = Path(sys.argv[1]).resolve()
    probe = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in list(_lib.SIGNATURES):
        if not hasattr(probe, name):
            _lib.SIGNATURES.pop(name)
from stableanimator_b200.arch import UNetConfig  # noqa: E402
from stableanimator_b200.synth import synth_inputs, synth_unet_state_dict  # noqa: E402
from stableanimator_b200.unet import UNetB200  # noqa: E402

cfg = UNetConfig.full()
unet = UNetB200(cfg, synth_unet_state_dict(cfg), "cuda")
B, F, h = 2, 16, 64
inp = synth_inputs(cfg, F, h, h)
ctx = unet.prepare_context(inp["encoder_hidden_states"][:B].cuda())
ati = inp["added_time_ids"][:B].cuda()
x = torch.randn(B * F, h, h, 8, device="cuda").bfloat16()
t = torch.tensor(1.0, device="cuda")
unet.forward_nhwc(x, t, ctx, ati, B, F, h, h)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    unet.forward_nhwc(x, t, ctx, ati, B, F, h, h)
ts = []
for _ in range(8):
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 5)
ts.sort()
print(f"{_lib.LIB_PATH.name:18s} fuse={unet.fuse_gn_stats}: median {ts[4]:.3f} ms  min {ts[0]:.3f}  max {ts[-1]:.3f}")
