"""CPU restatement of the CTC criterion that consumes the encoder's ``ctc_out``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline leg may import this module; the product path never does.

Follows
  * ``examples/speech_recognition/criterions/CTC_loss.py:31-74``   compute_ctc_uer
  * ``examples/speech_recognition/utils/wer_utils.py:80-96``       EditDistance.cost (not time mediated)
  * ``examples/speech_recognition/utils/wer_utils.py:98-139``      get_result (backtrace -> codes)
  * ``examples/speech_recognition/utils/wer_utils.py:141-202``     align (the DP and its tie-breaks)
  * ``examples/speech_recognition/criterions/CTC_loss.py:128-151`` the F.ctc_loss call (sum reduction,
    zero_infinity=True, blank = dictionary index of <ctc_blank>, on log_softmax of the fp32 logits:
    ``ctc_multi_loss.py:64-72``)

``F.ctc_loss`` itself is third-party arithmetic (PyTorch, unpinned in the reference's setup.py:136-144;
2.11.0 here): ``ctc_nll`` restates the published alpha recursion (Graves et al. 2006, eq. 6-8) in
float64 and is pinned against ``torch.nn.functional.ctc_loss`` in ``tests/test_oracle_criterion.py``;
``uer`` is pinned against the LIVE reference's ``compute_ctc_uer`` (golden file
``tests/golden/ctc_criterion.pt``, generator ``oracle/make_golden_criterion.py``).
"""
import math
from itertools import groupby

import numpy as np

COST_STEP = 3  # insertion / deletion   (wer_utils.py:93-94)
COST_SUB = 4   # substitution           (wer_utils.py:95-96)


def collapse(frame_labels, blank):
    """CTC_loss.py:50-58: dedup consecutive predictions, then drop blanks."""
    return [p for p, _ in groupby(frame_labels) if p != blank]


def align_errors(refs, hyps):
    """Number of non-match codes on the path EditDistance.align + get_result follow.

    The reference minimises the WEIGHTED cost (0/3/3/4) and backtracks one specific path chosen by
    strict '<' comparisons in the order diagonal, (i, j-1), (i-1, j)  (wer_utils.py:170-192); the
    error count is the length of that path minus its matches (CTC_loss.py:68-70), which is not in
    general the Levenshtein distance.  Both sequences empty: the reference returns NaN from align()
    and then fails on ``.codes``; defined here as 0 errors.
    """
    R, H = len(refs), len(hyps)
    if R == 0 and H == 0:
        return 0
    score = np.zeros((R + 1, H + 1), dtype=np.int64)
    back = np.zeros((R + 1, H + 1, 2), dtype=np.int64)
    for i in range(R + 1):
        for j in range(H + 1):
            if i == 0 and j == 0:
                continue
            if i == 0:
                score[i, j] = score[i, j - 1] + COST_STEP
                back[i, j] = (i, j - 1)
                continue
            if j == 0:
                score[i, j] = score[i - 1, j] + COST_STEP
                back[i, j] = (i - 1, j)
                continue
            best = score[i - 1, j - 1] + (0 if refs[i - 1] == hyps[j - 1] else COST_SUB)
            prev = (i - 1, j - 1)
            ins = score[i, j - 1] + COST_STEP
            if ins < best:
                best, prev = ins, (i, j - 1)
            dele = score[i - 1, j] + COST_STEP
            if dele < best:
                best, prev = dele, (i - 1, j)
            score[i, j] = best
            back[i, j] = prev
    errors = 0
    i, j = R, H
    while (i, j) != (0, 0):  # wer_utils.py:106-137
        pi, pj = back[i, j]
        if pi == i - 1 and pj == j - 1:
            if refs[i - 1] != hyps[j - 1]:
                errors += 1
        else:
            errors += 1
        i, j = int(pi), int(pj)
    return errors


def uer(frame_labels, in_lengths, targets, target_lengths, blank):
    """compute_ctc_uer given the per-frame arg-max labels [B][T] (CTC_loss.py:47-72).
    Returns (errors per utterance, collapsed prediction lengths, batch_errors, batch_total)."""
    errs, plens, total = [], [], 0
    for b in range(len(in_lengths)):
        pred = collapse([int(v) for v in frame_labels[b][: int(in_lengths[b])]], blank)
        tgt = [int(v) for v in targets[b][: int(target_lengths[b])]]
        errs.append(align_errors(pred, tgt))  # predicted tokens are passed as `refs` (:61-63)
        plens.append(len(pred))
        total += len(tgt)
    return errs, plens, sum(errs), total


def _lse(*xs):
    m = max(xs)
    if m == -math.inf:
        return -math.inf
    return m + math.log(sum(math.exp(x - m) for x in xs))


def ctc_nll(logprobs, in_len, target, blank):
    """-log p(target | x) for one utterance; logprobs [T, V] float64 (already normalised)."""
    U = len(target)
    S = 2 * U + 1
    ext = [blank if s % 2 == 0 else int(target[s // 2]) for s in range(S)]
    if in_len == 0:
        return 0.0 if U == 0 else math.inf
    alpha = [-math.inf] * S
    alpha[0] = float(logprobs[0, blank])
    if S > 1:
        alpha[1] = float(logprobs[0, ext[1]])
    for t in range(1, in_len):
        new = [-math.inf] * S
        for s in range(S):
            terms = [alpha[s]]
            if s >= 1:
                terms.append(alpha[s - 1])
            if s >= 3 and s % 2 == 1 and ext[s] != ext[s - 2]:
                terms.append(alpha[s - 2])
            new[s] = _lse(*terms) + float(logprobs[t, ext[s]])
        alpha = new
    ll = _lse(alpha[S - 1], alpha[S - 2]) if S > 1 else alpha[0]
    return -ll


def ctc_loss_sum(logits, in_lengths, targets, target_lengths, blank):
    """CTC_loss.py:143-151 on time-major logits [T, B, V] (any float dtype): log_softmax in float64,
    per-utterance nll, zero_infinity, sum.  Returns (per-utterance nll list, sum)."""
    x = np.asarray(logits, dtype=np.float64)
    m = x.max(axis=-1, keepdims=True)
    lp = x - (m + np.log(np.exp(x - m).sum(axis=-1, keepdims=True)))
    out = []
    for b in range(x.shape[1]):
        v = ctc_nll(lp[:, b], int(in_lengths[b]), [int(t) for t in targets[b][: int(target_lengths[b])]], blank)
        out.append(0.0 if math.isinf(v) or math.isnan(v) else v)
    return out, float(sum(out))
