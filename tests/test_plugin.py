"""The fairseq --user-dir plugin registers against the LIVE reference (CPU, build container only)."""
import argparse
import os
import sys

import pytest
import torch

from oracle import ref_loader as R

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not R.available(), reason="live reference not mounted")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "fbk-fairseq-st_b200", "fbkst_b200", "plugin")


@pytest.fixture(scope="module")
def plugin():
    R.load()
    from fairseq import utils
    utils.import_user_module(argparse.Namespace(user_dir=PLUGIN))  # what train.py --user-dir does
    return True


class Task:
    pass


def make_task():
    t = Task()
    t.source_dictionary = R.make_dictionary(60)
    t.target_dictionary = R.make_dictionary(70)
    return t


def test_registry(plugin):
    from fairseq.models import ARCH_MODEL_REGISTRY, MODEL_REGISTRY
    assert "conv_transformer_b200" in MODEL_REGISTRY
    for a in ("conv_transformer_b200", "conv_transformer_big_b200", "conv_transformer_big2_b200",
              "conv_transformer_giant_b200"):
        assert ARCH_MODEL_REGISTRY[a] is MODEL_REGISTRY["conv_transformer_b200"]


def test_build_model_matches_reference_state_dict(plugin):
    """Same args -> our model's state_dict keys/shapes == the reference model's, strict load works,
    and the encoder is a FairseqEncoder (fairseq_model.py:247)."""
    from fairseq.models import ARCH_CONFIG_REGISTRY, MODEL_REGISTRY, FairseqEncoder

    def args_for(arch):
        # the real CLI path: fairseq's own two-pass parser (fairseq/options.py:81-197)
        from fairseq import options
        parser = options.get_training_parser()
        return options.parse_args_and_arch(parser, [
            "/tmp/nodata", "--user-dir", PLUGIN, "--arch", arch,
            "--task", "speech_translation_with_transcription", "--criterion", "ctc_multi_loss",
            "--underlying-criterion", "label_smoothed_cross_entropy", "--ctc-encoder-layer", "2",
            "--ctc-compress-out", "--ctc-compress-strategy", "avg", "--no-attn-2d",
            "--distance-penalty", "log", "--input-feat-per-channel", "40", "--encoder-layers", "3",
            "--decoder-layers", "1", "--max-tokens", "1000", "--skip-normalization"])
    task = make_task()
    ours = MODEL_REGISTRY["conv_transformer_b200"].build_model(args_for("conv_transformer_big2_b200"), task)
    ref = MODEL_REGISTRY["conv_transformer"].build_model(args_for("conv_transformer_big2"), task)
    assert isinstance(ours.encoder, FairseqEncoder)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert set(sd_ref.keys()) == set(sd_ours.keys())
    for k, v in sd_ref.items():
        assert tuple(v.shape) == tuple(sd_ours[k].shape), k
    ours.load_state_dict(sd_ref, strict=True)
    assert ours.encoder.embed_dim == 512 and ours.encoder.heads == 8 and ours.encoder.log_penalty


def test_unsupported_flags_raise(plugin):
    from fairseq.models import ARCH_CONFIG_REGISTRY, MODEL_REGISTRY
    a = argparse.Namespace(distance_penalty="log", ctc_compress_out=False, criterion="x",
                           input_feat_per_channel=40, encoder_layerdrop=0.0, decoder_layerdrop=0.0)
    ARCH_CONFIG_REGISTRY["conv_transformer_b200"](a)  # attn_2d defaults to True without --no-attn-2d
    with pytest.raises(NotImplementedError, match="no-attn-2d"):
        MODEL_REGISTRY["conv_transformer_b200"].build_model(a, make_task())
