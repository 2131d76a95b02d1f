"""Error metrics shared by the parity tests.

north_star asks for "per-kernel max relative error <= 2e-3 in bf16"; SURVEY.md section 8c defines it
as max |out - ref| relative to the tensor's max-abs, ref = fp32 oracle on the same bf16-rounded
inputs.  A bf16 OUTPUT carries a rounding error of up to 2^-8 = 3.9e-3 of its own magnitude (8-bit
significand), so for the elements near the max a plain max-norm of 2e-3 is unattainable for ANY
bf16-valued result -- including the reference's own bf16 path.  We therefore bound:
  * excess   = max(|out - ref| - 2^-8 |ref|, 0) / max|ref|    <= 2e-3   (the survey's metric, net of
                                                                         the output's own bf16 rounding)
  * rel_l2   = ||out - ref||_2 / ||ref||_2                    <= 2^-8   (sanity bound: pure bf16 output
               rounding alone gives ~1.6e-3; a kernel with one internal bf16 operand, e.g. the
               attention probabilities P, ~2.3e-3 -- the same as any bf16 flash attention)
For fp32 outputs ``excess`` degenerates to the plain max-norm relative error and both use ``tol``.
"""
import torch

TOL = 2e-3
L2_TOL_BF16 = 2.0 ** -8


def rel_l2(out, ref):
    out, ref = out.double(), ref.double()
    return ((out - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def excess(out, ref):
    o, r = out.double(), ref.double()
    quant = 0.0 if out.dtype == torch.float32 else 2.0 ** -8
    e = ((o - r).abs() - quant * r.abs()).clamp_min(0)
    return (e.max() / r.abs().max().clamp_min(1e-30)).item()


def assert_close(out, ref, tol=TOL, what=""):
    assert out.shape == ref.shape, f"{what}: shape {tuple(out.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(out.float()).all(), f"{what}: non-finite output"
    l2, ex = rel_l2(out, ref), excess(out, ref)
    l2_tol = tol if out.dtype == torch.float32 else max(tol, L2_TOL_BF16)
    assert ex <= tol and l2 <= l2_tol, f"{what}: excess_max={ex:.3e} (tol {tol}) rel_l2={l2:.3e} (tol {l2_tol:.1e})"
    return l2, ex
