"""Gradient parity of the whole differentiable encoder (scope row T) against torch.autograd of the CPU oracle
(the restatement of ConvolutionalTransformerEncoder.forward pinned to the live reference; its gradients are
pinned to the live reference's in tests/test_oracle.py::test_oracle_gradients_vs_live_reference).

Deterministic mode: eval() with gradients enabled (BatchNorm running statistics, no dropout) -- the reference's
training mode draws dropout masks from torch's generator (and the conv dropout p = max(dropout, .1) cannot be
switched off), so bitwise training-mode parity does not exist; training mode is covered by kernel-level tests
(tests/test_gpu_train_ops.py) and an end-to-end optimisation test here.

Tolerance: 2e-2 of each gradient tensor's max magnitude (bf16 operands; north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import build_encoder, rel_err  # noqa: E402
from oracle import encoder_oracle as O  # noqa: E402  (checker only)

TOL = 2e-2


def oracle_grads(sd, cfg, x, lens, hook, r_out_fn, r_ctc_fn):
    leaf = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running" not in k and "_float_tensor" not in k:
            v.requires_grad_(True)
        leaf[k] = v
    ref = O.encoder_forward(leaf, cfg, x, lens, ctc_logits_hook=hook)
    r_out, r_ctc = r_out_fn(ref), r_ctc_fn(ref)
    loss = (ref["encoder_out"] * r_out).sum()
    if ref["ctc_out"] is not None:
        loss = loss + (ref["ctc_out"] * r_ctc).sum()
    loss.backward()
    return ref, {k: v.grad for k, v in leaf.items() if v.requires_grad}, r_out, r_ctc


def run_case(cfg, lens_in, seed, feat=40):
    sd = O.init_state_dict(cfg, seed=seed)
    x, lens = O.synthetic_batch(lens_in, feat, seed=seed + 100)
    T = max(lens_in)
    L = ((T + 1) // 2 + 1) // 2
    hook = None
    if cfg.get("ctc_layer", 0) > 0:
        labels = O.synthetic_ctc_bump(L, len(lens_in), cfg["vocab"], seed=seed + 7)
        hook = O.bump_hook(labels, 30.0)
    g = torch.Generator().manual_seed(seed + 1)

    def r_out_fn(ref):  # upstream gradient of encoder_out: zero at padded positions (what the decoder produces)
        r = torch.randn(ref["encoder_out"].shape, generator=g)
        for b, n in enumerate(ref["src_lengths"].tolist()):
            r[n:, b] = 0
        return r

    def r_ctc_fn(ref):  # upstream gradient of the CTC logits: zero beyond each input length (F.ctc_loss)
        if ref["ctc_out"] is None:
            return None
        r = torch.randn(ref["ctc_out"].shape, generator=g) * 0.05
        sub = [((n + 1) // 2 + 1) // 2 for n in lens_in]
        for b, n in enumerate(sub):
            r[n:, b] = 0
        return r

    ref, grads, r_out, r_ctc = oracle_grads(sd, cfg, x, lens, hook, r_out_fn, r_ctc_fn)
    enc = build_encoder(cfg, sd)  # eval(): running statistics, no dropout
    if hook is not None:
        dev_hook = O.bump_hook(labels.cuda(), 30.0)
        enc.ctc_fc.register_forward_hook(lambda m, i, o: dev_hook(o))
    for p in enc.parameters():
        p.requires_grad_(True)
    out = enc(x.cuda(), lens.cuda(), return_all_hiddens=True)
    assert out.src_lengths.cpu().tolist() == ref["src_lengths"].tolist()
    assert out.encoder_out.requires_grad
    assert rel_err(out.encoder_out.detach(), ref["encoder_out"].detach()) < TOL
    loss = (out.encoder_out * r_out.cuda()).sum()
    if r_ctc is not None:
        assert rel_err(out.ctc_out.detach(), ref["ctc_out"].detach()) < TOL
        loss = loss + (out.ctc_out * r_ctc.cuda()).sum()
    loss.backward()
    worst = {}
    for name, p in enc.named_parameters():
        assert p.grad is not None, name
        assert torch.isfinite(p.grad).all(), name
        e = rel_err(p.grad, grads[name])
        worst[name] = e
    bad = {k: round(v, 4) for k, v in worst.items() if v >= TOL}
    assert not bad, bad
    return worst


def test_gradients_tiny_log_penalty_ctc():
    cfg = dict(embed_dim=128, ffn_dim=256, heads=2, layers=3, conv_channels=64, feat_dim=40, vocab=64,
               distance_penalty="log", ctc_layer=2, ctc_strategy="weighted")
    run_case(cfg, [97, 64, 30], seed=3)


def test_gradients_mha_layout_no_compression():
    cfg = dict(embed_dim=256, ffn_dim=512, heads=4, layers=2, conv_channels=64, feat_dim=40, vocab=50,
               distance_penalty=None, ctc_layer=0, ctc_strategy="avg")
    run_case(cfg, [201, 160, 77, 40], seed=5)


@pytest.mark.parametrize("strategy", ["avg", "softmax"])
def test_gradients_big2_shape(strategy):
    """EACL'21 model shape (d512 h8 ffn2048, log penalty, V=1005) at reduced depth, ragged batch with odd conv
    lengths (the convolutions' gradients flow through padded frames exactly as in the reference)."""
    cfg = dict(embed_dim=512, ffn_dim=2048, heads=8, layers=3, conv_channels=64, feat_dim=40, vocab=1005,
               distance_penalty="log", ctc_layer=2, ctc_strategy=strategy)
    run_case(cfg, [601, 598, 411, 203], seed=7)


def test_gradients_giant_shape():
    cfg = dict(embed_dim=1024, ffn_dim=4096, heads=16, layers=2, conv_channels=128, feat_dim=80, vocab=305,
               distance_penalty="log", ctc_layer=1, ctc_strategy="avg")
    run_case(cfg, [1210, 517], seed=9, feat=80)


def test_training_mode_optimises():
    """train(): BatchNorm batch statistics + every dropout site + backward.  A few SGD steps on one batch must
    reduce a simple loss, running statistics must move, and two forwards with the same seed state must agree."""
    cfg = dict(embed_dim=128, ffn_dim=256, heads=2, layers=2, conv_channels=64, feat_dim=40, vocab=64,
               distance_penalty="log", ctc_layer=0, ctc_strategy="avg", dropout=0.1)
    sd = O.init_state_dict(cfg, seed=11)
    enc = build_encoder(cfg, sd).train()
    x, lens = O.synthetic_batch([120, 99, 64], 40, seed=12)
    x, lens = x.cuda(), lens.cuda()
    target = torch.randn(30, 3, 128, device="cuda") * 0.1
    rm0 = enc.bn[0].running_mean.clone()
    opt = torch.optim.SGD(enc.parameters(), lr=0.05)
    losses = []
    for step in range(12):
        torch.manual_seed(100 + step)
        out = enc(x, lens, return_all_hiddens=True)
        loss = ((out.encoder_out - target) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses
    assert not torch.equal(enc.bn[0].running_mean, rm0)
    assert int(enc.bn[0].num_batches_tracked) == 12
    torch.manual_seed(5)
    a = enc(x, lens).encoder_out.detach().clone()
    torch.manual_seed(5)
    b = enc(x, lens).encoder_out.detach()
    assert torch.equal(a, b)  # same seed -> same dropout masks
    torch.manual_seed(6)
    c = enc(x, lens).encoder_out.detach()
    assert not torch.equal(a, c)
    enc.eval()
    with torch.no_grad():
        e1 = enc(x, lens).encoder_out
    assert torch.isfinite(e1).all()
