"""Pin the criterion oracle: golden vectors made by the LIVE reference's
compute_ctc_uer + the criterion's F.ctc_loss call, hand-derived known answers for the alignment
tie-breaks, torch's F.ctc_loss on random cases, and (when mounted) the live reference itself."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import ctc_criterion_oracle as C
from oracle import ref_loader as R


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "ctc_criterion.pt"), weights_only=False)


def test_collapse_kat():
    assert C.collapse([5, 5, 9, 9, 2, 2, 2, 9, 2, 3, 3], 9) == [5, 2, 2, 3]  # CTC_loss.py:50-58
    assert C.collapse([9, 9, 9], 9) == []
    assert C.collapse([], 9) == []


def test_align_errors_kat():
    """Weighted costs 0/3/3/4 with the reference's tie-breaks: the count is the number of non-match
    steps on ONE specific path, not the Levenshtein distance."""
    assert C.align_errors([1, 2, 3], [1, 2, 3]) == 0
    assert C.align_errors([1, 2, 3], [1, 9, 3]) == 1          # one substitution (cost 4 < 3+3)
    assert C.align_errors([1, 2, 3], []) == 3
    assert C.align_errors([], [4, 5]) == 2
    assert C.align_errors([], []) == 0                          # reference: undefined (raises)
    assert C.align_errors([1, 2], [2, 1]) == 2                 # step+match+step (6) beats sub+sub (8): 2 errors either way
    assert C.align_errors([1, 2, 3, 4], [2, 3, 4]) == 1
    assert C.align_errors([7, 7, 7], [7]) == 2


def test_uer_matches_golden(golden):
    for name, c in golden.items():
        lab = c["frame_labels"].t().tolist()  # [B][T]
        errs, plens, tot_e, tot_n = C.uer(lab, c["in_lengths"].tolist(), c["targets"].tolist(),
                                          c["target_lengths"].tolist(), c["blank"])
        for b, want in enumerate(c["ref_errors"]):
            if want is not None:
                assert errs[b] == want, (name, b, errs[b], want)
            else:
                assert errs[b] == 0
        assert tot_n == c["ref_total"]


def test_ctc_loss_matches_golden(golden):
    for name, c in golden.items():
        per, tot = C.ctc_loss_sum(c["logits"].numpy(), c["in_lengths"].tolist(), c["targets"].tolist(),
                                  c["target_lengths"].tolist(), c["blank"])
        ref = c["ref_nll"].tolist()
        for b in range(len(per)):
            assert abs(per[b] - ref[b]) <= 1e-4 * max(1.0, abs(ref[b])), (name, b, per[b], ref[b])
        assert abs(tot - c["ref_loss"]) <= 1e-4 * max(1.0, abs(c["ref_loss"]))


def test_ctc_loss_vs_torch_random():
    g = torch.Generator().manual_seed(5)
    for T, B, V, U in ((20, 3, 6, 5), (9, 2, 4, 6), (33, 4, 11, 1)):
        logits = torch.randn(T, B, V, generator=g)
        il = torch.randint(1, T + 1, (B,), generator=g)
        tl = torch.randint(0, U + 1, (B,), generator=g)
        tg = torch.randint(0, V - 1, (B, U), generator=g)
        flat = torch.cat([tg[b, : tl[b]] for b in range(B)])
        ref = F.ctc_loss(F.log_softmax(logits, -1), flat, il, tl, blank=V - 1, reduction="none",
                         zero_infinity=True)
        per, _ = C.ctc_loss_sum(logits.numpy(), il.tolist(), tg.tolist(), tl.tolist(), V - 1)
        assert torch.allclose(torch.tensor(per, dtype=torch.float64), ref.double(), rtol=1e-4, atol=1e-4), (per, ref)


@pytest.mark.reference
@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
def test_uer_vs_live_reference():
    R.load()
    from examples.speech_recognition.criterions.CTC_loss import compute_ctc_uer
    g = torch.Generator().manual_seed(11)
    for trial in range(6):
        T, B, V = 30, 4, 6
        logits = torch.randn(T, B, V, generator=g) * 3
        il = torch.randint(1, T + 1, (B,), generator=g)
        tl = torch.randint(1, 12, (B,), generator=g)
        tg = torch.randint(0, V - 1, (B, 12), generator=g)
        lp = F.log_softmax(logits, -1).transpose(0, 1)
        e, n = compute_ctc_uer(lp, tg, il, tl, V - 1)
        _, _, oe, on = C.uer(lp.argmax(-1).tolist(), il.tolist(), tg.tolist(), tl.tolist(), V - 1)
        assert (oe, on) == (int(e), int(n))
