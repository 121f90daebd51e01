"""GPU parity of the CTC criterion device path, through the C ABI.

  * golden vectors produced by the LIVE reference (compute_ctc_uer, F.ctc_loss as the criterion calls it)
  * the CPU oracle on seeded inputs (bf16 and fp32 logits, ragged lengths, edge cases)
  * full-size cfg2 shapes through size-independent properties

Errors / totals / collapsed lengths are integers: bit-exact.  nll is fp32 arithmetic over fp32 (or
bf16-stored) logits: 1e-4 relative (the sum of exponentials uses ex2.approx)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ctc_criterion_oracle as C  # noqa: E402  (checker only)


def dev():
    return torch.device("cuda:0")


def run_device(logits, il, tg, tl, blank):
    from fbkst_b200 import criterion
    loss, nll, errors, totals = criterion.ctc_loss_and_uer(logits.to(dev()), il.to(dev()), tg.to(dev()),
                                                           tl.to(dev()), blank)
    torch.cuda.synchronize()
    return loss.item(), nll.cpu().double(), errors.cpu().tolist(), totals.cpu().tolist()


def test_criterion_golden(golden_dir):
    golden = torch.load(os.path.join(golden_dir, "ctc_criterion.pt"), weights_only=False)
    for name, c in golden.items():
        loss, nll, errors, totals = run_device(c["logits"], c["in_lengths"], c["targets"],
                                               c["target_lengths"], c["blank"])
        want = [0 if e is None else e for e in c["ref_errors"]]
        assert errors == want, (name, errors, want)
        assert totals == [sum(want), c["ref_total"]], (name, totals)
        assert torch.allclose(nll, c["ref_nll"], rtol=1e-4, atol=1e-4), (name, nll, c["ref_nll"])
        assert abs(loss - c["ref_loss"]) <= 1e-4 * max(1.0, abs(c["ref_loss"])), (name, loss, c["ref_loss"])


def test_reference_signature_compute_ctc_uer(golden_dir):
    """compute_ctc_uer(logprobs N x T x D, targets, input_lengths, target_lengths, blank) as the
    criterion calls it (CTC_loss.py:153-156): transposed view of the T x N x D log-probabilities."""
    from fbkst_b200 import criterion
    c = torch.load(os.path.join(golden_dir, "ctc_criterion.pt"), weights_only=False)["long"]
    lp = F.log_softmax(c["logits"].to(dev()), dim=-1).transpose(0, 1)
    e, n = criterion.compute_ctc_uer(lp, c["targets"].to(dev()), c["in_lengths"].to(dev()),
                                     c["target_lengths"].to(dev()), c["blank"])
    assert (e, n) == (float(sum(c["ref_errors"])), float(c["ref_total"]))
    # a genuinely batch-first contiguous tensor is accepted as well
    e2, n2 = criterion.compute_ctc_uer(lp.contiguous(), c["targets"].to(dev()), c["in_lengths"].to(dev()),
                                       c["target_lengths"].to(dev()), c["blank"])
    assert (e2, n2) == (e, n)
    with pytest.raises(ValueError):
        criterion.compute_ctc_uer(lp.cpu(), c["targets"], c["in_lengths"], c["target_lengths"], c["blank"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T,B,V,U,seed", [(50, 5, 17, 12, 0), (90, 7, 33, 40, 1), (30, 3, 8, 0, 2),
                                          (200, 4, 1005, 70, 3), (64, 33, 40, 300, 4)])
def test_criterion_vs_oracle(T, B, V, U, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(T, B, V, generator=g) * 2).to(dtype)
    # run structure so that collapsing matters; margin large enough that bf16 keeps the arg-max
    path = torch.randint(0, V, (T // 3 + 1, B), generator=g).repeat_interleave(3, dim=0)[:T]
    logits.scatter_add_(2, path.unsqueeze(-1), torch.full((T, B, 1), 8.0, dtype=dtype))
    il = torch.randint(1, T + 1, (B,), generator=g)
    il[0] = T
    tl = torch.randint(0, U + 1, (B,), generator=g)
    tg = torch.randint(0, V - 1, (B, max(U, 1)), generator=g)
    blank = V - 1
    loss, nll, errors, totals = run_device(logits, il, tg, tl, blank)
    lab = logits.float().argmax(-1).t().tolist()
    oe, _, osum, on = C.uer(lab, il.tolist(), tg.tolist(), tl.tolist(), blank)
    assert errors == oe
    assert totals == [osum, on]
    per, tot = C.ctc_loss_sum(logits.float().numpy(), il.tolist(), tg.tolist(), tl.tolist(), blank)
    per = torch.tensor(per, dtype=torch.float64)
    assert torch.allclose(nll, per, rtol=1e-4, atol=1e-3), (nll, per)
    assert abs(loss - tot) <= 1e-4 * max(1.0, abs(tot))


def test_criterion_full_size_properties():
    """cfg2 shape (T'=375, B=64, V=8005, bf16 logits with the row pitch the ctc_fc GEMM writes):
    (1) a target equal to the collapsed arg-max path has 0 errors and dropping its last k tokens gives
    exactly k errors (pure deletions are the unique cheapest alignment of a prefix... of distinct
    tokens), (2) totals are the sums, (3) nll matches torch's F.ctc_loss on the same device tensors
    (library call used as a checker), (4) run-to-run identical."""
    from fbkst_b200 import criterion, ops
    T, B, V, blank = 375, 64, 8005, 8004
    g = torch.Generator().manual_seed(9)
    buf = torch.zeros(T * B, 8008, dtype=torch.bfloat16)
    buf[:, :V] = (torch.randn(T * B, V, generator=g) * 0.5).to(torch.bfloat16)
    lens = torch.randint(200, T + 1, (B,), generator=g)
    lens[0] = T
    # every run gets a fresh, never repeated label so that the collapsed path has distinct tokens
    path = torch.full((T, B), blank, dtype=torch.long)
    targets, tls = [], []
    for b in range(B):
        t, nxt, seq = 0, b * 100, []
        while t < int(lens[b]):
            n = int(torch.randint(1, 6, (1,), generator=g))
            if torch.rand(1, generator=g).item() < 0.5:
                path[t:t + n, b] = nxt
                seq.append(nxt)
                nxt += 1
            t += n
        k = b % 4
        seq = seq[: len(seq) - k] if k else seq
        targets.append(seq)
        tls.append(len(seq))
    rows = torch.arange(T * B)
    buf[rows, path.reshape(-1)] += 12.0
    U = max(tls)
    tg = torch.ones(B, U, dtype=torch.long)
    for b, s in enumerate(targets):
        tg[b, : len(s)] = torch.tensor(s)
    logits = buf.to(dev()).view(T, B, 8008)[:, :, :V]  # column-narrowed view, pitch 8008
    il, tl, tgd = lens.to(dev()), torch.tensor(tls).to(dev()), tg.to(dev())
    loss, nll, errors, totals = criterion.ctc_loss_and_uer(logits, il, tgd, tl, blank)
    loss2, nll2, errors2, totals2 = criterion.ctc_loss_and_uer(logits, il, tgd, tl, blank)
    torch.cuda.synchronize()
    assert errors.cpu().tolist() == [b % 4 for b in range(B)]
    assert totals.cpu().tolist() == [sum(b % 4 for b in range(B)), sum(tls)]
    assert torch.equal(nll, nll2) and torch.equal(loss, loss2) and torch.equal(errors, errors2)
    lp = F.log_softmax(logits.float(), dim=-1)
    ref = F.ctc_loss(lp, tgd, il, tl, blank=blank, reduction="none", zero_infinity=True)
    assert torch.allclose(nll.double(), ref.double(), rtol=1e-4, atol=1e-3), (nll, ref)
    assert abs(loss.item() - ref.sum().item()) <= 1e-4 * max(1.0, abs(ref.sum().item()))
    assert ops.LAUNCHES > 0


# ------------------------------------------------------------------ loss backward (N1, second half)
@pytest.mark.parametrize("T,B,V,U", [(50, 4, 30, 12), (375, 6, 1005, 40), (120, 3, 8005, 25)])
def test_ctc_loss_backward_vs_torch(T, B, V, U):
    """d loss / d logits of the device CTC loss == autograd of F.ctc_loss(log_softmax(logits), ..., "sum",
    zero_infinity=True) (criterions/CTC_loss.py:143-151), incl. repeated labels, ragged input / target
    lengths and one infeasible utterance (target longer than its input: zero loss, zero gradient)."""
    from fbkst_b200 import criterion as C
    g = torch.Generator().manual_seed(T + V)
    logits = torch.randn(T, B, V, generator=g).cuda()
    in_len = torch.tensor([T] + [max(3, T - 7 * b) for b in range(1, B)], dtype=torch.long)
    in_len[-1] = 4  # infeasible with a longer target
    tgt_len = torch.tensor([min(U, max(1, int(in_len[b]) // 3)) for b in range(B)], dtype=torch.long)
    tgt_len[-1] = min(U, 9)
    blank = V - 1
    targets = torch.randint(4, V - 1, (B, U), generator=g)
    targets[0, 1] = targets[0, 0]  # repeated label (needs a blank in between)
    targets[0, 3] = targets[0, 0]  # same label again later: one writer per column
    for b in range(B):
        targets[b, tgt_len[b]:] = 1
    mask = (torch.arange(T)[None, :] >= in_len[:, None]).cuda()  # B x T, True = padding
    x = logits.clone().requires_grad_(True)
    lp = torch.log_softmax(x, -1)
    flat = torch.cat([targets[b, :tgt_len[b]] for b in range(B)]).cuda()
    ref = torch.nn.functional.ctc_loss(lp, flat, in_len.cuda(), tgt_len.cuda(), blank=blank, reduction="sum",
                                       zero_infinity=True)
    (ref * 0.7).backward()
    y = logits.clone().requires_grad_(True)
    loss, totals, il = C.ctc_loss_train(y, mask.t(), targets.cuda(), tgt_len.cuda(), blank)
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    (loss * 0.7).backward()
    err = (y.grad - x.grad).abs().max().item() / x.grad.abs().max().item()
    assert err < 1e-3, err  # fp32 on both sides (different summation orders in the alpha / beta recursions)
    assert y.grad[:, -1].abs().max() == 0  # infeasible utterance
    assert y.grad[int(in_len[1]):, 1].abs().max() == 0  # frames beyond the input length
    y2 = logits.clone().requires_grad_(True)
    l2, _, _ = C.ctc_loss_train(y2, mask.t(), targets.cuda(), tgt_len.cuda(), blank)
    (l2 * 0.7).backward()
    assert torch.equal(y2.grad, y.grad)  # one writer per element: run-to-run identical
