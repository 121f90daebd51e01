"""Floating-point parity AT FULL DEPTH on the configurations BASELINE.json names.

The bench's own cfg2 model (11L d512 h8 ffn2048, log penalty, V=8005, CTC compression @8, 64 x 1500 x 40;
the weights, inputs and label plan `bench.py` uses) against the CPU oracle restatement of
`ConvolutionalTransformerEncoder.forward` (conv_transformer.py:195-276), all three
`--ctc-compress-strategy` values, eager AND CUDA-graph mode; a true cfg3 batch (ragged 200..3000
frames, odd conv lengths: SURVEY F5) at 11 layers; a 12-layer cfg5 batch (d1024 h16 C128 F80).

Asserted: compressed lengths and padding masks bit-exact; encoder outputs under BOTH readings of the
north_star's "2e-2 relative under bf16" (tests/helpers.py): max-normalised and element-wise.
The measured numbers are appended to gpurun_out/parity_full.json (copied to profiles/ by hand).
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

import bench  # noqa: E402  (the headline configuration lives there)
from helpers import TOL_BF16, parity_report  # noqa: E402
from oracle import encoder_oracle as O  # noqa: E402  (checker only)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(name, rep):
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_full.json"), "a") as f:
            f.write(json.dumps(dict(test=name, **rep)) + "\n")
    except OSError:
        pass


def _bench_encoder(model):
    """Exactly bench.run_ours' model construction."""
    from fbkst_b200.config import build_encoder
    torch.manual_seed(0)
    enc = build_encoder(model, None, device="cpu")
    bench.randomise_norm_stats(enc, 1)
    return enc.cuda().eval()


def _device_bump(enc, plan, margin=bench.CTC_MARGIN):
    state = dict(plan=plan.cuda())

    def bump(m, i, o):  # out of place: capturable, leaves the GEMM output buffer alone
        p = state["plan"]
        return o.scatter_add(2, p.unsqueeze(-1), torch.full_like(o[..., :1], margin))
    return enc.ctc_fc.register_forward_hook(bump), state


def _compare(name, out, ref, cols=None):
    """lengths / masks exact, floats under both criteria.  ``cols``: batch columns of ``out`` that
    the reference batch holds (default: all)."""
    B_ref = ref["encoder_out"].shape[1]
    cols = list(range(B_ref)) if cols is None else cols
    nl = ref["src_lengths"].tolist()
    assert out.src_lengths.cpu()[cols].tolist() == nl, name
    L2 = max(nl)
    eo = out.encoder_out[:L2, cols].float().cpu()
    if ref["encoder_padding_mask"] is not None:
        assert out.encoder_padding_mask is not None
        assert torch.equal(out.encoder_padding_mask.cpu()[cols][:, :L2], ref["encoder_padding_mask"]), name
    rep = parity_report(eo, ref["encoder_out"], nl)
    _record(name, rep)
    assert rep["max_rel"] < TOL_BF16, (name, rep)
    assert rep["elementwise"] < TOL_BF16, (name, rep)
    assert torch.isfinite(out.encoder_out).all()
    return rep


@pytest.mark.parametrize("strategy", ["avg", "weighted", "softmax"])
def test_cfg2_full_depth_vs_oracle(strategy):
    """BASELINE configs[1] as the bench runs it; the oracle computes the first 8 of the 64 utterances
    (all 1500 frames long: nothing is padded, so an utterance's result does not depend on its batch)."""
    cfgb = bench.CONFIGS["cfg2"]
    model = dict(cfgb["model"], ctc_strategy=strategy)
    lengths = cfgb["lengths"]
    B, T, Fd = len(lengths), max(lengths), model["feat_dim"]
    L = ((T + 1) // 2 + 1) // 2
    enc = _bench_encoder(model)
    plan = bench.label_plan(L, B, model["vocab"], seed=7)
    _, state = _device_bump(enc, plan)
    x, lens = bench.make_batch(lengths, Fd, 1234)

    from fbkst_b200 import ops
    xn = ops.cmvn(x.cuda(), lens.to(torch.int32).cuda())
    out_eager = enc(xn, lens.cuda())  # device lengths: the fairseq call path (utils.move_to_cuda)
    enc.use_cuda_graph = True
    out_graph = enc(xn, lens)  # host lengths: the bench's call path
    torch.cuda.synchronize()
    assert torch.equal(out_eager.src_lengths, out_graph.src_lengths)
    assert torch.equal(out_eager.encoder_out, out_graph.encoder_out), "graph replay != eager launches"

    n_ref = 8
    sd = {k: v.detach().float().cpu() for k, v in enc.state_dict().items()}
    xr = torch.zeros(n_ref, T, Fd)
    for b in range(n_ref):  # data/fbank_dataset.py:44-45: CMVN per utterance
        xr[b, : lengths[b]] = O.cmvn(x[b, : lengths[b]])
    hook = O.bump_hook(plan[:, :n_ref], bench.CTC_MARGIN)
    ref = O.encoder_forward(sd, model, xr, lens[:n_ref], ctc_logits_hook=hook)
    # the 8 reference utterances are compared over THEIR compressed extent
    for name, out in (("cfg2/%s/eager" % strategy, out_eager), ("cfg2/%s/graph" % strategy, out_graph)):
        nl = ref["src_lengths"].tolist()
        assert out.src_lengths.cpu()[:n_ref].tolist() == nl
        rep = parity_report(out.encoder_out[:, :n_ref].float().cpu(), ref["encoder_out"], nl)
        _record(name, rep)
        assert rep["max_rel"] < TOL_BF16 and rep["elementwise"] < TOL_BF16, (name, rep)
        if ref["encoder_padding_mask"] is not None:
            L2 = ref["encoder_out"].shape[0]
            assert torch.equal(out.encoder_padding_mask.cpu()[:n_ref, :L2], ref["encoder_padding_mask"])
    # CTC logits themselves (fp32 on both sides, bf16 GEMM operands on ours): informative bound
    lg = out_eager.ctc_out[:, :n_ref].float().cpu()
    rep = parity_report(lg, ref["ctc_out"], [L] * n_ref)
    _record("cfg2/%s/ctc_logits" % strategy, rep)
    assert rep["max_rel"] < TOL_BF16, rep


@pytest.mark.parametrize("strategy", ["weighted", "softmax"])
def test_cfg3_full_depth_ragged_vs_oracle(strategy):
    """BASELINE configs[2]: the cfg2 model at its full 11 layers on a ragged 200..3000-frame batch with
    odd conv lengths (identical batch composition on both sides: SURVEY F5)."""
    model = dict(bench.CONFIGS["cfg3"]["model"], ctc_strategy=strategy)
    lens_in = [3000, 2999, 2750, 2501, 2222, 2001, 1777, 1502, 1333, 1001, 999, 801, 602, 403, 250, 201]
    B, T, Fd = len(lens_in), max(lens_in), model["feat_dim"]
    L = ((T + 1) // 2 + 1) // 2
    enc = _bench_encoder(model)
    plan = bench.label_plan(L, B, model["vocab"], seed=11)
    _device_bump(enc, plan)
    x, lens = bench.make_batch(lens_in, Fd, 4321)
    xn = torch.zeros_like(x)
    for b, n in enumerate(lens_in):
        xn[b, :n] = O.cmvn(x[b, :n])
    sd = {k: v.detach().float().cpu() for k, v in enc.state_dict().items()}
    ref = O.encoder_forward(sd, model, xn, lens, ctc_logits_hook=O.bump_hook(plan, bench.CTC_MARGIN))
    from fbkst_b200 import ops
    xd = ops.cmvn(x.cuda(), lens.to(torch.int32).cuda())
    out = enc(xd, lens.cuda())
    _compare("cfg3/%s/eager" % strategy, out, ref)
    enc.use_cuda_graph = True
    out_g = enc(xd, lens.cuda())
    assert torch.equal(out_g.encoder_out, out.encoder_out) and torch.equal(out_g.src_lengths, out.src_lengths)


def test_cfg5_full_depth_vs_oracle():
    """BASELINE configs[4]: 12 layers d1024 h16 ffn4096, 128 conv channels, 80-dim fbank, 3000..6000-frame
    utterances, compression at layer 8."""
    model = dict(bench.CONFIGS["cfg5"]["model"])
    lens_in = [6000, 4001, 3000]
    B, T, Fd = len(lens_in), max(lens_in), model["feat_dim"]
    L = ((T + 1) // 2 + 1) // 2
    enc = _bench_encoder(model)
    plan = bench.label_plan(L, B, model["vocab"], seed=13)
    _device_bump(enc, plan)
    x, lens = bench.make_batch(lens_in, Fd, 99)
    xn = torch.zeros_like(x)
    for b, n in enumerate(lens_in):
        xn[b, :n] = O.cmvn(x[b, :n])
    sd = {k: v.detach().float().cpu() for k, v in enc.state_dict().items()}
    ref = O.encoder_forward(sd, model, xn, lens, ctc_logits_hook=O.bump_hook(plan, bench.CTC_MARGIN))
    from fbkst_b200 import ops
    xd = ops.cmvn(x.cuda(), lens.to(torch.int32).cuda())
    out = enc(xd, lens.cuda())
    _compare("cfg5/avg/eager", out, ref)


def test_unbumped_logits_same_logits_contract():
    """No logit injection: random-init `ctc_fc`, so the top-2 gap of many frames is tiny.  The
    contract (north_star): labels / segment boundaries / lengths are bit-exact WHEN BOTH SIDES ARE FED
    THE SAME CTC LOGITS -- the oracle's softmax+argmax+groupby (conv_transformer.py:282-287) is run on
    the fp32 logits our forward returns, and must give our compressed lengths."""
    model = dict(embed_dim=512, ffn_dim=2048, heads=8, layers=3, conv_channels=64, feat_dim=40,
                 vocab=8005, distance_penalty="log", ctc_layer=2, ctc_strategy="avg")
    lens_in = [1500, 1203, 997, 640]
    enc = _bench_encoder(model)
    x, lens = bench.make_batch(lens_in, 40, 77)
    out = enc(x.cuda(), lens.cuda())
    assert out.ctc_out.dtype == torch.float32
    sub = torch.tensor([((n + 1) // 2 + 1) // 2 for n in lens_in])
    segs = O.ctc_segments(out.ctc_out.float().cpu().contiguous(), sub)
    assert out.src_lengths.cpu().tolist() == [len(s) for s in segs]
    # with random logits nearly every frame is its own run: the compression must not be degenerate
    assert max(len(s) for s in segs) > 300


@pytest.mark.parametrize("strategy", ["avg", "weighted"])
def test_fused_ctc_epilogue_equals_hook_path(strategy):
    """The bench's logit injection through the fused ctc_fc + arg-max epilogue (enc.ctc_logit_bump) gives
    exactly the compressed lengths of the forward-hook path and the same outputs (same kernels otherwise)."""
    cfgb = bench.CONFIGS["cfg2"]
    model = dict(cfgb["model"], ctc_strategy=strategy, layers=9)
    lens_in = [1500, 1500, 1203, 997, 640, 333]
    B, T, Fd = len(lens_in), 1500, 40
    L = 375
    enc = _bench_encoder(model)
    plan = bench.label_plan(L, B, model["vocab"], seed=5)
    x, lens = bench.make_batch(lens_in, Fd, 31)
    handle, _ = _device_bump(enc, plan)
    o_hook = enc(x.cuda(), lens.cuda())
    handle.remove()
    enc.ctc_logit_bump = (plan.to(torch.int32).cuda().contiguous(), bench.CTC_MARGIN)
    o_fused = enc(x.cuda(), lens.cuda())
    assert torch.equal(o_fused.src_lengths, o_hook.src_lengths)
    nl = o_hook.src_lengths.tolist()
    rep = parity_report(o_fused.encoder_out.float().cpu(), o_hook.encoder_out.float().cpu(), nl)
    assert rep["max_rel"] < 1e-3, rep  # weighted: probabilities from exp2f partials vs the arg-max kernel's
    assert torch.allclose(o_fused.ctc_out[:, :, :], o_hook.ctc_out, atol=1e-4)
    enc.use_cuda_graph = True
    o_graph = enc(x.cuda(), lens)
    assert torch.equal(o_graph.encoder_out, o_fused.encoder_out)
