"""The real drop-in path on a B200: fairseq's own ``--user-dir`` import, its
two-pass argument parser, ``build_model`` of ``conv_transformer_big2_b200``, a STRICT load of the
reference model's state_dict, and fairseq's ``SequenceGenerator`` driving our encoder through
``forward_torchscript`` / ``reorder_encoder_out`` (fairseq/sequence_generator.py:193-198, 703-709)
next to the UNMODIFIED reference model on the same GPU.

Needs the reference package: ``/root/reference`` (build container) or its copy ``baseline/_ref``
(``baseline/make_ref.py``; travels to the GPU box).  Skipped when neither exists.
"""
import argparse
import os
import warnings

import pytest
import torch

from oracle import ref_loader as R  # checker side: imports the reference with its shims

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="reference package (baseline/_ref) not present")]

from helpers import TOL_BF16, parity_report  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "fbk-fairseq-st_b200", "fbkst_b200", "plugin")


class _Task:
    pass


def _args(arch, extra=()):
    from fairseq import options
    parser = options.get_training_parser()
    return options.parse_args_and_arch(parser, [
        "/tmp/nodata", "--user-dir", PLUGIN, "--arch", arch,
        "--task", "speech_translation_with_transcription", "--criterion", "ctc_multi_loss",
        "--underlying-criterion", "label_smoothed_cross_entropy", "--ctc-encoder-layer", "2",
        "--ctc-compress-out", "--ctc-compress-strategy", "avg", "--no-attn-2d",
        "--distance-penalty", "log", "--input-feat-per-channel", "40", "--encoder-layers", "3",
        "--decoder-layers", "2", "--max-tokens", "1000", "--skip-normalization"] + list(extra))


@pytest.fixture(scope="module")
def models():
    warnings.simplefilter("ignore")
    R.load()
    from fairseq import utils
    utils.import_user_module(argparse.Namespace(user_dir=PLUGIN))  # what train.py/generate.py --user-dir do
    from fairseq.models import MODEL_REGISTRY
    task = _Task()
    task.source_dictionary = R.make_dictionary(120)
    task.target_dictionary = R.make_dictionary(90)
    torch.manual_seed(3)
    ref = MODEL_REGISTRY["conv_transformer"].build_model(_args("conv_transformer_big2"), task)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():  # BatchNorm away from the identity, non-zero biases
        for bn in ref.encoder.bn:
            bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.1)
            bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
        for n, p in ref.named_parameters():
            if n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    ours = MODEL_REGISTRY["conv_transformer_b200"].build_model(_args("conv_transformer_big2_b200"), task)
    ours.load_state_dict(ref.state_dict(), strict=True)  # the checkpoint boundary
    ours_x = MODEL_REGISTRY["conv_transformer_b200"].build_model(
        _args("conv_transformer_big2_b200", ["--b200-cross-attention"]), task)
    ours_x.load_state_dict(ref.state_dict(), strict=True)
    return task, ref.cuda().eval(), ours.cuda().eval(), ours_x.cuda().eval()


def _sample(lens_in, seed):
    g = torch.Generator().manual_seed(seed)
    B, T = len(lens_in), max(lens_in)
    x = torch.randn(B, T, 40, generator=g)
    for b, n in enumerate(lens_in):
        x[b, n:] = 0
    from fairseq import utils
    return utils.move_to_cuda(dict(net_input=dict(
        src_tokens=x, src_lengths=torch.tensor(lens_in, dtype=torch.long),
        prev_output_tokens=torch.zeros(B, 1, dtype=torch.long))))  # dropped by forward_non_torchscript


def _bump(models_, L, B, vocab, seed):
    import bench
    plan = bench.label_plan(L, B, vocab, seed=seed).cuda()
    handles = []
    for m in models_:
        handles.append(m.encoder.ctc_fc.register_forward_hook(
            lambda mod, i, o: o.scatter_add(2, plan[: o.shape[0]].unsqueeze(-1),
                                            torch.full_like(o[..., :1], 30.0))))
    return handles


def test_generate_through_the_plugin(models):
    task, ref, ours, ours_x = models
    lens_in = [801, 640, 523, 402, 200]
    B, L = len(lens_in), (max(lens_in) + 3) // 4
    handles = _bump([ref, ours, ours_x], L, B, len(task.source_dictionary), seed=21)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # fp32 reference, not TF32
    try:
        sample = _sample(lens_in, 5)
        from fairseq.sequence_generator import SequenceGenerator
        with torch.no_grad():
            # (1) the encoder exactly as SequenceGenerator calls it (device-resident int64 lengths)
            eo_ref = ref.encoder.forward_torchscript(sample["net_input"])
            eo = ours.encoder.forward_torchscript(sample["net_input"])
            assert type(eo).__name__ == type(eo_ref).__name__  # fairseq's own NamedTuple type
            nl = eo_ref.src_lengths.tolist()
            assert eo.src_lengths.tolist() == nl
            assert eo.encoder_out.shape == eo_ref.encoder_out.shape
            assert torch.equal(eo.encoder_padding_mask, eo_ref.encoder_padding_mask)
            rep = parity_report(eo.encoder_out.float().cpu(), eo_ref.encoder_out.float().cpu(), nl)
            assert rep["max_rel"] < TOL_BF16 and rep["elementwise"] < TOL_BF16, rep

            # (2) beam search with the reference model, then with ours (stock and device cross-attention)
            gen_ref = SequenceGenerator([ref], task.target_dictionary, beam_size=4, max_len_b=12)
            hyp_ref = gen_ref.generate([ref], sample)
            agree = {}
            for name, m in (("stock-decoder", ours), ("b200-cross-attention", ours_x)):
                gen = SequenceGenerator([m], task.target_dictionary, beam_size=4, max_len_b=12)
                hyp = gen.generate([m], sample)
                assert len(hyp) == B and all(len(h) == 4 for h in hyp)
                same = sum(int(h[0]["tokens"].tolist() == r[0]["tokens"].tolist()) for h, r in zip(hyp, hyp_ref))
                agree[name] = same
                # random-init decoders have near-flat distributions: a bf16-sized perturbation may flip a
                # beam choice, so token identity is asserted for most, not all, utterances ...
                assert same >= B - 2, (name, same)

            # (3) ... and the decoder's log-probabilities are compared teacher-forced on the reference's
            # own best hypotheses (no search in the loop)
            pad = task.target_dictionary.pad()
            eos = task.target_dictionary.eos()
            toks = [r[0]["tokens"] for r in hyp_ref]
            U = max(len(t) for t in toks)
            prev = torch.full((B, U), pad, dtype=torch.long, device="cuda")
            for b, t in enumerate(toks):
                prev[b, 0] = eos
                prev[b, 1:len(t)] = t[:-1]
            lp_ref = ref.get_normalized_probs(ref.decoder(prev, encoder_out=eo_ref), log_probs=True)
            for m in (ours, ours_x):
                lp = m.get_normalized_probs(m.decoder(prev, encoder_out=m.encoder.forward_torchscript(
                    sample["net_input"])), log_probs=True)
                for b, t in enumerate(toks):
                    d = (lp[b, :len(t)] - lp_ref[b, :len(t)]).abs().max().item()
                    assert d < 5e-2, (b, d)
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
        for h in handles:
            h.remove()


class _CritTask(_Task):
    def build_criterion(self, args):
        from fairseq import criterions
        return criterions.build_criterion(args, self)


def test_train_step_through_the_criterion():
    """fairseq's training call path (tasks/speech_recognition.py:234-263): criterion(model, sample) ->
    loss.backward(), for the reference model with ``ctc_multi_loss`` and for ours with
    ``--arch conv_transformer_big2_b200 --criterion ctc_multi_loss_b200`` (encoder forward/backward and the
    CTC loss on the sm_100a kernels; decoder and label-smoothed CE are fairseq's in both).  Deterministic mode
    (eval() + grad: no dropout, BatchNorm running statistics).  Loss values within 2e-2; every parameter's
    gradient has cosine >= 0.99 with the reference's (see tests/test_gpu_train.py for why max-normalised
    errors of ReLU-gated tensors are noise under bf16)."""
    warnings.simplefilter("ignore")
    R.load()
    from fairseq import criterions, utils
    utils.import_user_module(argparse.Namespace(user_dir=PLUGIN))
    from fairseq.models import MODEL_REGISTRY
    task = _CritTask()
    task.source_dictionary = R.make_dictionary(120)
    task.target_dictionary = R.make_dictionary(90)
    extra = ["--label-smoothing", "0.1"]
    a_ref = _args("conv_transformer_big2", extra)
    a_our = _args("conv_transformer_big2_b200", extra)
    a_our.criterion = "ctc_multi_loss_b200"
    torch.manual_seed(7)
    ref = MODEL_REGISTRY["conv_transformer"].build_model(a_ref, task)
    ours = MODEL_REGISTRY["conv_transformer_b200"].build_model(a_our, task)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref, ours = ref.cuda().eval(), ours.cuda().eval()
    crit_ref = criterions.build_criterion(a_ref, task).cuda()
    crit_our = criterions.build_criterion(a_our, task).cuda()
    assert type(crit_our).__name__ == "B200CTCMultiLoss"
    lens_in = [801, 640, 523, 402]
    B, L = len(lens_in), (max(lens_in) + 3) // 4
    handles = _bump([ref, ours], L, B, len(task.source_dictionary), seed=23)
    g = torch.Generator().manual_seed(1)
    s = _sample(lens_in, 9)
    U, U1 = 11, 9
    tgt = torch.randint(4, 89, (B, U), generator=g)
    tgt[:, -1] = task.target_dictionary.eos()
    prev = torch.cat([torch.full((B, 1), task.target_dictionary.eos()), tgt[:, :-1]], 1)
    tr = torch.randint(4, 118, (B, U1), generator=g)
    trl = torch.tensor([9, 8, 7, 5])
    for b in range(B):
        tr[b, trl[b]:] = task.source_dictionary.pad()
    sample = utils.move_to_cuda(dict(
        net_input=dict(src_tokens=s["net_input"]["src_tokens"].cpu(), src_lengths=s["net_input"]["src_lengths"].cpu(),
                       prev_output_tokens=prev),
        target=tgt, transcript_target=tr, transcript_target_lengths=trl, ntokens=int(B * U), nsentences=B))
    try:
        with R.grad_shims():
            loss_ref, ss_ref, log_ref = crit_ref(ref, sample)
            loss_ref.backward()
        loss_our, ss_our, log_our = crit_our(ours, sample)
        loss_our.backward()
        assert ss_ref == ss_our
        assert abs(log_our["ctc_loss"] - log_ref["ctc_loss"]) <= 2e-2 * abs(log_ref["ctc_loss"])
        assert abs(loss_our.item() - loss_ref.item()) <= 2e-2 * abs(loss_ref.item())
        assert log_our["ctc_errors"] == log_ref["ctc_errors"] and log_our["ctc_total"] == log_ref["ctc_total"]
        assert log_our["nframes"] == log_ref["nframes"]
        pr = dict(ref.named_parameters())
        low = {}
        for name, p in ours.named_parameters():
            gr = pr[name].grad
            if gr is None or p.grad is None:
                assert gr is None and p.grad is None, name
                continue
            a, r = p.grad.double().flatten(), gr.double().flatten()
            if r.abs().max() < 1e-9 or name.endswith("k_proj.bias"):
                continue
            c = (torch.dot(a, r) / (a.norm() * r.norm()).clamp_min(1e-30)).item()
            if c < 0.99:
                low[name] = round(c, 4)
        assert not low, low
    finally:
        for h in handles:
            h.remove()
