"""Timeline of one tile of sa_ff_fused (CTA 0): builds a private copy of the library from ff_fused.cu with -DSA_FF_TRACE,
runs the L0 problem and prints, per hidden chunk, the clock64() stamps of the hand-off points relative to S(0)."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from stableanimator_b200 import _lib, build, ops  # noqa: E402

out = ROOT / "gpurun_out" / "libsa_fftrace.so"
out.parent.mkdir(exist_ok=True)
csrc = ROOT / "stableanimator_b200" / "csrc"
cmd = [build._nvcc(), *build.NVCC_FLAGS, "-DSA_FF_TRACE", "-shared", "-o", str(out), str(csrc / "ff_fused.cu"), str(csrc / "common.cu")]
subprocess.run(cmd, check=True)
lib = C.CDLL(str(out))
lib.sa_ff_fused.argtypes = [C.POINTER(_lib.FfArgs), C.c_void_p]
lib.sa_ff_set_trace.argtypes = [C.c_void_p]

m, c = 131072, 320
x = torch.randn(m, c, device="cuda").bfloat16()
w1 = (torch.randn(8 * c, c, device="cuda") * c ** -0.5).bfloat16()
b1 = torch.randn(8 * c, device="cuda")
w2 = (torch.randn(c, 4 * c, device="cuda") * (4 * c) ** -0.5).bfloat16()
b2 = torch.randn(c, device="cuda")
g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
wi, bi = ops.interleave_geglu(w1, b1)
o = torch.empty(m, c, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(4 * 64, device="cuda", dtype=torch.int64)
assert lib.sa_ff_set_trace(trace.data_ptr()) == 0
a = _lib.FfArgs()
a.x, a.ldx, a.m, a.c = x.data_ptr(), c, m, c
a.gamma, a.beta, a.eps = g.data_ptr(), b.data_ptr(), 1e-5
a.w1, a.b1, a.w2, a.b2 = wi.data_ptr(), bi.data_ptr(), w2.data_ptr(), b2.data_ptr()
a.out, a.ldc, a.rows_per_group = o.data_ptr(), c, 1
for _ in range(3):
    assert lib.sa_ff_fused(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
torch.cuda.synchronize()
t = trace.cpu().reshape(4, 64)
t0 = int(t[0, 0])
print("chunk  S_issued  S_seen  P_written  O_issued   (cycles since S(0) issue; seen/written: set leader warp)")
for i in range(20):
    print(f"{i:5d} {int(t[0, i]) - t0:9d} {int(t[2, i]) - t0:7d} {int(t[3, i]) - t0:10d} {int(t[1, i]) - t0:9d}")
