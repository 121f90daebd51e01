"""Shared helpers for the GPU parity tests (test infrastructure)."""
from fbkst_b200.config import DictStub, build_encoder, make_args  # noqa: F401


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
