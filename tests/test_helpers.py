"""The parity helpers themselves (CPU): a non-finite value on either side must read as an infinite error.
Round 2 found the bench's parity key printing 0.0 for outputs full of NaN (max() drops a NaN)."""
import math
import os
import sys

import torch

from helpers import elementwise_err, parity_report, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_rel_err_and_elementwise_err_flag_non_finite_values():
    ref = torch.randn(7, 5)
    good = ref + 1e-3 * torch.randn(7, 5)
    assert rel_err(good, ref) < 1e-2 and elementwise_err(good, ref) < 1e-1
    for poison in (float("nan"), float("inf"), float("-inf")):
        bad = good.clone()
        bad[3, 2] = poison
        assert math.isinf(rel_err(bad, ref)) and math.isinf(elementwise_err(bad, ref))
        assert math.isinf(rel_err(good, bad)) and math.isinf(elementwise_err(good, bad))


def test_parity_report_does_not_hide_a_nan_utterance():
    ref = torch.randn(6, 3, 4)
    out = ref.clone()
    out[:, 1] = float("nan")  # one whole utterance (the failure mode of DESIGN.md 4e)
    rep = parity_report(out, ref, [6, 6, 4])
    assert math.isinf(rep["max_rel"]) and math.isinf(rep["elementwise"])
    out[:, 1] = ref[:, 1]
    out[5, 2] = float("nan")  # beyond utterance 2's length: not compared
    rep = parity_report(out, ref, [6, 6, 4])
    assert rep["max_rel"] == 0.0 and rep["elementwise"] == 0.0


def test_bench_parity_numbers_reports_non_finite_counts():
    import bench
    ref = torch.randn(6, 3, 4)
    out = ref + 1e-4
    par = bench.parity_numbers(out, ref, [6, 6, 4])
    assert par["max_rel"] < 1e-2 and par["nonfinite_ours"] == 0 and par["nonfinite_reference"] == 0
    out[2, 0, 1] = float("nan")
    par = bench.parity_numbers(out, ref, [6, 6, 4])
    assert math.isinf(par["max_rel"]) and math.isinf(par["elementwise"]) and par["nonfinite_ours"] == 1
