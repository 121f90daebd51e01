"""Regenerate ``tests/golden/ctc_criterion.pt`` from the LIVE reference -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_criterion

Runs the reference's own ``compute_ctc_uer`` (criterions/CTC_loss.py:31-74) and the exact
``F.ctc_loss`` call of the criterion (CTC_loss.py:143-151) on seeded synthetic cases and stores inputs
and outputs.  Needs ``/root/reference`` (build container only); the fixture travels, this script's
import of the reference does not.
"""
import os
import warnings

import torch
import torch.nn.functional as F

from . import ref_loader

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                      "ctc_criterion.pt")
PAD = 1  # fairseq Dictionary.pad()


def make_case(seed, T, B, V, blank, in_lengths, targets, margin=6.0, run_mean=2.5, blank_p=0.45):
    """Logits whose arg-max path has CTC-like runs; targets are given or derived from a
    corrupted copy of the collapsed path so that the alignment has matches, substitutions,
    insertions and deletions."""
    g = torch.Generator().manual_seed(seed)
    logits = 0.5 * torch.randn(T, B, V, generator=g)
    path = torch.zeros(T, B, dtype=torch.long)
    for b in range(B):
        t = 0
        while t < T:
            n = 1 + int(torch.empty(1).geometric_(1.0 / run_mean, generator=g).item()) - 1
            n = max(1, n)
            lab = blank if torch.rand(1, generator=g).item() < blank_p else int(
                torch.randint(0, V, (1,), generator=g).item())
            path[t:t + n, b] = lab
            t += n
    logits.scatter_add_(2, path.unsqueeze(-1), torch.full((T, B, 1), margin))
    if targets is None:
        targets = []
        for b in range(B):
            seq, prev = [], None
            for v in path[: in_lengths[b], b].tolist():
                if v != prev and v != blank:
                    seq.append(v)
                prev = v
            out = []
            for v in seq:  # corrupt: drop / substitute / duplicate / insert
                r = torch.rand(1, generator=g).item()
                if r < 0.12:
                    continue
                if r < 0.24:
                    v = int(torch.randint(0, V, (1,), generator=g).item())
                    v = v if v != blank else (blank + 1) % V
                out.append(v)
                if r > 0.88:
                    out.append(v)
                if 0.80 < r <= 0.88:
                    w = int(torch.randint(0, V, (1,), generator=g).item())
                    out.append(w if w != blank else (blank + 1) % V)
            targets.append(out)
    U = max(1, max(len(t) for t in targets))
    tgt = torch.full((B, U), PAD, dtype=torch.long)
    for b, t in enumerate(targets):
        tgt[b, : len(t)] = torch.tensor(t, dtype=torch.long)
    return {"logits": logits, "in_lengths": torch.tensor(in_lengths, dtype=torch.long), "targets": tgt,
            "target_lengths": torch.tensor([len(t) for t in targets], dtype=torch.long), "blank": blank,
            "pad": PAD}


def run_reference(case):
    ref_loader.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from examples.speech_recognition.criterions.CTC_loss import compute_ctc_uer
    lprobs = F.log_softmax(case["logits"].float(), dim=-1)  # ctc_multi_loss.py:64-67, T x B x V
    tl, il = case["target_lengths"], case["in_lengths"]
    flat = case["targets"].masked_select(
        torch.arange(case["targets"].shape[1])[None, :] < tl[:, None])  # CTC_loss.py:140-141
    per = F.ctc_loss(lprobs, flat, il, tl, blank=case["blank"], reduction="none", zero_infinity=True)
    tot = F.ctc_loss(lprobs, flat, il, tl, blank=case["blank"], reduction="sum", zero_infinity=True)
    per_utt = []
    for b in range(lprobs.shape[1]):  # per-utterance error counts through the same reference function
        if int(tl[b]) == 0 and all(v == case["blank"] for v in lprobs[: il[b], b].argmax(-1).tolist()):
            per_utt.append(None)  # the reference raises (align() returns NaN): undefined there
            continue
        e, n = compute_ctc_uer(lprobs.transpose(0, 1)[b:b + 1], case["targets"][b:b + 1], il[b:b + 1],
                               tl[b:b + 1], case["blank"])
        assert n == int(tl[b])
        per_utt.append(int(e))
    case["ref_errors"] = per_utt
    case["ref_total"] = int(tl.sum())
    case["ref_nll"] = per.double()
    case["ref_loss"] = float(tot)
    case["frame_labels"] = lprobs.argmax(-1)  # T x B
    return case


def cases():
    out = {}
    out["runs"] = make_case(0, 48, 6, 14, 13, [48, 45, 40, 33, 20, 9], None)
    out["blank0"] = make_case(1, 37, 4, 9, 0, [37, 30, 22, 5], None)  # blank need not be the last index
    # hand-made: repeated target labels (need a blank between), target longer than the input can
    # emit (nll = inf -> 0), empty target with non-blank predictions, single frame
    out["edge"] = make_case(2, 12, 5, 7, 6, [12, 12, 3, 10, 1],
                            [[1, 1, 2, 2, 2], [3, 4, 3, 4, 3, 4, 5], [1, 1, 1], [], [2]], blank_p=0.3)
    out["long"] = make_case(3, 160, 3, 40, 39, [160, 131, 77], None, run_mean=3.0)
    return out


def main():
    data = {k: run_reference(c) for k, c in cases().items()}
    torch.save(data, GOLDEN)
    for k, c in data.items():
        print(k, "errors", c["ref_errors"], "total", c["ref_total"], "loss %.4f" % c["ref_loss"])


if __name__ == "__main__":
    main()
