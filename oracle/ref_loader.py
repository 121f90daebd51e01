"""Import the LIVE reference (``/root/reference``) with in-process shims.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works where the reference
tree is mounted (the build container) or where ``baseline/make_ref.py`` has made its copy
``baseline/_ref/`` (which travels to the GPU box).  The reference tree is never
edited; all fixes are monkeypatches applied in this process:

  shim 1  ``np.float/np.int/np.bool`` aliases (fairseq/data/indexed_dataset.py:89
          uses the numpy aliases removed in numpy 1.24)               [SURVEY F1]
  shim 2  ``EncoderOut._field_types`` (conv_transformer.py:30 reads the
          NamedTuple attribute removed in Python 3.9)                 [SURVEY F1]
  shim 3  ``torch.Tensor.cuda`` no-op while running on CPU
          (local_attention.py:132 hard-codes ``.cuda()``)             [SURVEY F2]
  shim 4  ``LocalAttention.in_proj_qkv`` clones its chunks when gradients are on
          (in-place ``q *= scaling`` on a view, local_attention.py:98)  [SURVEY F11]
"""
import argparse
import contextlib
import os
import sys
import warnings

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """FBKST_REFERENCE_ROOT, else the mounted tree (build container), else the copy made by
    ``baseline/make_ref.py`` (git-ignored; the only form of the reference that reaches the GPU box)."""
    env = os.environ.get("FBKST_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "examples", "speech_recognition")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_ct = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "examples", "speech_recognition"))


def load():
    """Return the reference module ``examples.speech_recognition.models.conv_transformer``."""
    global _ct
    if _ct is not None:
        return _ct
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    import numpy as np

    for name, typ in (("float", float), ("int", int), ("bool", bool)):
        if name not in np.__dict__:
            setattr(np, name, typ)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from fairseq.models import transformer as T

        if not hasattr(T.EncoderOut, "_field_types"):
            T.EncoderOut._field_types = T.EncoderOut.__annotations__
        import examples.speech_recognition.models.conv_transformer as ct
    _ct = ct
    return ct


@contextlib.contextmanager
def cpu_cuda_noop():
    """Shim 3: make ``Tensor.cuda`` a no-op so LocalAttention runs on CPU."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


@contextlib.contextmanager
def grad_shims():
    """Shim 4: with gradients enabled, ``q *= self.scaling`` (local_attention.py:98) modifies in
    place a view returned by ``chunk`` (:152-153), which torch >= 1.x autograd rejects.  Cloning the three
    chunks is the same arithmetic.  Needed only for gradient parity (eval() + grad) on the reference."""
    load()
    from examples.speech_recognition.modules.local_attention import LocalAttention
    orig = LocalAttention.in_proj_qkv
    LocalAttention.in_proj_qkv = lambda self, query: tuple(t.clone() for t in orig(self, query))
    try:
        yield
    finally:
        LocalAttention.in_proj_qkv = orig


def make_dictionary(vocab: int):
    """fairseq Dictionary with ``vocab`` entries in total, the last being
    ``<ctc_blank>`` (tasks/speech_translation_ctc.py:43-44)."""
    load()
    from fairseq.data import Dictionary

    d = Dictionary()
    for i in range(vocab - len(d) - 1):
        d.add_symbol("w%d" % i)
    d.add_symbol("<ctc_blank>")
    assert len(d) == vocab, (len(d), vocab)
    return d


def make_args(cfg: dict) -> argparse.Namespace:
    """Namespace the reference ctor reads (conv_transformer.py:134-193), filled by
    the reference's own ``base_architecture`` defaults (:429-466)."""
    ct = load()
    a = argparse.Namespace()
    a.encoder_embed_dim = cfg["embed_dim"]
    a.encoder_ffn_embed_dim = cfg["ffn_dim"]
    a.encoder_attention_heads = cfg["heads"]
    a.encoder_layers = cfg["layers"]
    a.encoder_convolutions = "[(%d, 3, 3)] * 2" % cfg.get("conv_channels", 64)
    a.input_feat_per_channel = cfg["feat_dim"]
    a.distance_penalty = "log" if cfg.get("distance_penalty", "log") == "log" else False
    a.no_attn_2d = True
    a.ctc_compress_out = cfg.get("ctc_layer", 0) > 0
    a.ctc_compress_strategy = cfg.get("ctc_strategy", "avg")
    a.ctc_encoder_layer = cfg.get("ctc_layer", 0)
    a.criterion = "ctc_multi_loss"
    a.max_source_positions = 100000
    a.max_target_positions = 100000
    a.encoder_layerdrop = 0.0
    a.dropout = cfg.get("dropout", 0.1)
    ct.base_architecture(a)
    return a


def build_reference_encoder(cfg: dict, seed: int = 0, randomize_bn: bool = True):
    """Reference ``ConvolutionalTransformerEncoder`` with the reference's own
    initialisers under ``torch.manual_seed(seed)``; BatchNorm running stats are
    randomised so that BN is not the identity."""
    ct = load()
    args = make_args(cfg)
    torch.manual_seed(seed)
    enc = ct.ConvolutionalTransformerEncoder(
        args, make_dictionary(cfg["vocab"]), audio_features=cfg["feat_dim"])
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for bn in enc.bn:
            bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.1)
            bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
            bn.weight.data.copy_(1.0 + 0.1 * torch.randn(bn.weight.shape, generator=g))
            bn.bias.data.copy_(0.1 * torch.randn(bn.bias.shape, generator=g))
        # non-trivial biases / LN affine so that every epilogue term is exercised
        for name, p in enc.named_parameters():
            if name.endswith(".bias") and "bn." not in name:
                p.data.copy_(0.05 * torch.randn(p.shape, generator=g))
            if "layer_norm.weight" in name:
                p.data.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
    enc.eval()
    return enc


def run_reference_encoder(enc, src_tokens, src_lengths, return_all_hiddens=False):
    """Reference forward on CPU: eval mode, no grad, with shim 3."""
    with torch.no_grad(), cpu_cuda_noop(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return enc(src_tokens, src_lengths, return_all_hiddens=return_all_hiddens)
