"""Shared helpers for the GPU parity tests (test infrastructure)."""
import torch

from fbkst_b200.config import DictStub, build_encoder, make_args  # noqa: F401

# BASELINE.json north_star: floating-point encoder outputs within 2e-2 under bf16.  Two readings
# of "relative" are asserted side by side:
#   max-normalised : max|a-b| / max|ref|                              < TOL
#   element-wise   : |a-b| <= TOL*|ref| + TOL*rms(ref)  for EVERY element (allclose with the
#                    absolute floor tied to the tensor's own scale, not to its largest value)
TOL_BF16 = 2e-2


def rel_err(a, b):
    """max|a-b| / max|ref|; inf when either side holds a non-finite value (a NaN must never compare as small:
    Python's max(x, nan) keeps x and torch's max() propagates NaN only by luck of the call site)."""
    a, b = a.double().cpu(), b.double().cpu()
    if not (bool(torch.isfinite(a).all()) and bool(torch.isfinite(b).all())):
        return float("inf")
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def elementwise_err(a, b):
    """max over elements of |a-b| / (|ref| + rms(ref)): the element-wise criterion holds iff the
    returned value is <= TOL."""
    a, b = a.double().cpu(), b.double().cpu()
    if not (bool(torch.isfinite(a).all()) and bool(torch.isfinite(b).all())):
        return float("inf")
    rms = b.pow(2).mean().sqrt().clamp_min(1e-12)
    return ((a - b).abs() / (b.abs() + rms)).max().item()


def parity_report(out, ref_out, lengths):
    """Both criteria over the valid positions of a T x B x D output (per utterance b only the first
    lengths[b] time steps are compared: padding rows are compared separately, exactly)."""
    worst_max, worst_el = 0.0, 0.0
    for b, n in enumerate(lengths):
        if n == 0:
            continue
        worst_max = max(worst_max, rel_err(out[:n, b], ref_out[:n, b]))
        worst_el = max(worst_el, elementwise_err(out[:n, b], ref_out[:n, b]))
    return dict(max_rel=worst_max, elementwise=worst_el)
