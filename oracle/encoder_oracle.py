"""CPU restatement of the reference ST-encoder forward -- TEST INFRASTRUCTURE ONLY.

A plain PyTorch (CPU, fp32 unless told otherwise) restatement of what
``ConvolutionalTransformerEncoder.forward`` computes in eval mode, written as
functions over a ``state_dict`` so that it needs neither fairseq nor
``/root/reference`` at run time (the GPU box has neither).  Every function cites
the reference file:line it follows (paths relative to ``/root/reference``; ``ST/``
abbreviates ``examples/speech_recognition/``).

Pinning: the reference's own tests hold no vector for this path ("parity
unpinned" by its suite).  This restatement is pinned against the
LIVE reference in the build container (tests/test_oracle_vs_reference.py and the
committed ``tests/golden/*.pt`` made by ``oracle/make_golden.py``) and against
the hand-derivable CTC-compression KAT of SURVEY.md 3.5.

The product path (``fbkst_b200``) never imports this module.
"""
import math
from itertools import groupby
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- a1
def cmvn(features: Tensor) -> Tensor:
    """Per-utterance mean/variance normalisation of one T x F fbank matrix.

    ST/data/data_utils.py:9-24 -- mean and UNBIASED variance over time; if any
    feature's variance is < 1e-8 the eps is added to sqrt(var) for ALL features.
    """
    if features.dim() != 2:
        raise ValueError("We expect the input feature to be 2-D tensor")
    mean = features.mean(0)
    var = features.var(0)
    eps = 1e-8
    if bool((var < eps).any()):
        inv = 1.0 / (torch.sqrt(var) + eps)
    else:
        inv = 1.0 / torch.sqrt(var)
    return (features - mean) * inv


# --------------------------------------------------------------------------- a5
def create_mask(lengths: Tensor) -> Optional[Tensor]:
    """ST/models/conv_transformer.py:293-300 -- B x max(len) bool, True = pad;
    ``None`` when nothing is padded."""
    max_len = int(lengths.max())
    mask = torch.arange(max_len).unsqueeze(0) >= lengths.unsqueeze(1)
    if not bool(mask.any()):
        return None
    return mask


# --------------------------------------------------------------------------- a4
def sinusoidal_table(num: int, dim: int) -> Tensor:
    """fairseq/modules/sinusoidal_positional_embedding.py:36-58 with
    padding_idx=0: row p = [sin(p*w_k) | cos(p*w_k)], w_k = exp(-k ln(1e4)/(dim/2-1)),
    row 0 zeroed."""
    half = dim // 2
    w = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    ang = torch.arange(num, dtype=torch.float).unsqueeze(1) * w.unsqueeze(0)
    tab = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if dim % 2 == 1:
        tab = torch.cat([tab, torch.zeros(num, 1)], dim=1)
    tab[0, :] = 0
    return tab


def positional_embedding(lengths: Tensor, dim: int) -> Tensor:
    """ST/modules/positional_embedding_audio.py:20-26 + fairseq/utils.py:192-202:
    position t+1 for t < len_b, 0 (zero row) for padding.  Returns B x L x dim."""
    max_len = int(lengths.max())
    t = torch.arange(max_len).unsqueeze(0)
    pos = torch.where(t < lengths.unsqueeze(1), t + 1, torch.zeros_like(t))
    return sinusoidal_table(max_len + 1, dim)[pos]


# --------------------------------------------------------------------------- a2
def _relu(x: Tensor, masks: Optional[dict], key: str) -> Tensor:
    """ReLU, or -- gradient-parity experiments only -- the same activation PATTERN as another
    implementation: ``masks[key]`` (0/1, x's shape) replaces ``x > 0`` in both the value and the gradient.
    A bf16 implementation flips the sign of pre-activations that are ~0; under random upstream gradients each
    flipped unit moves a weight gradient by an O(1) term, which is noise, not a defect (tests/test_gpu_train.py)."""
    if masks is None or key not in masks:
        return F.relu(x)
    return x * masks[key].to(x.dtype)


def conv_subsample(sd: Dict[str, Tensor], src_tokens: Tensor, src_lengths: Tensor,
                   bn_eps: float = 1e-5, relu_masks: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """ST/models/conv_transformer.py:202-214 in eval mode: for each of the two
    convs: Conv2d(k3,s2,p1)+bias -> ReLU -> BatchNorm2d(running stats) ->
    lengths = ceil(lengths/2).  Runs over the zero-padded batch with NO masking
    (SURVEY F5).  Returns (B x C x T' x F'', lengths')."""
    x = src_tokens.unsqueeze(1)
    lengths = src_lengths
    for i in range(2):
        x = F.conv2d(x, sd["convolutions.%d.weight" % i], sd["convolutions.%d.bias" % i],
                     stride=2, padding=1)
        x = _relu(x, relu_masks, "conv%d" % i)
        x = F.batch_norm(x, sd["bn.%d.running_mean" % i], sd["bn.%d.running_var" % i],
                         sd["bn.%d.weight" % i], sd["bn.%d.bias" % i], False, 0.0, bn_eps)
        lengths = torch.ceil(lengths.float() / 2).long()
    return x, lengths


# --------------------------------------------------------------------------- a3
def flatten_fc3(sd: Dict[str, Tensor], x: Tensor, relu_masks: Optional[dict] = None) -> Tensor:
    """ST/models/conv_transformer.py:225-227 -- (B,C,T',F'') -> (T',B,C*F'') with
    channel-major flatten (index c*F''+f), then ReLU(fc3(.))."""
    b, c, t, f = x.shape
    x = x.transpose(1, 2).reshape(b, t, c * f).transpose(0, 1)
    return _relu(F.linear(x, sd["fc3.weight"], sd["fc3.bias"]), relu_masks, "fc3")


# ----------------------------------------------------------------------- a7/a7'
def self_attention(sd: Dict[str, Tensor], prefix: str, x: Tensor, mask: Optional[Tensor],
                   heads: int, log_penalty: bool) -> Tensor:
    """ST/modules/local_attention.py:49-150 (log penalty: ST/modules/
    conv_transformer_layer.py:22-27) or, without penalty, the math of
    fairseq/modules/multihead_attention.py:145-177.  x is L x B x D."""
    L, B, D = x.shape
    hd = D // heads
    if prefix + "in_proj_weight" in sd:
        qkv = F.linear(x, sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)
    else:
        q = F.linear(x, sd[prefix + "q_proj.weight"], sd[prefix + "q_proj.bias"])
        k = F.linear(x, sd[prefix + "k_proj.weight"], sd[prefix + "k_proj.bias"])
        v = F.linear(x, sd[prefix + "v_proj.weight"], sd[prefix + "v_proj.bias"])
    q = q * (hd ** -0.5)
    q = q.contiguous().view(L, B * heads, hd).transpose(0, 1)
    k = k.contiguous().view(L, B * heads, hd).transpose(0, 1)
    v = v.contiguous().view(L, B * heads, hd).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))
    if mask is not None:
        s = s.view(B, heads, L, L).float().masked_fill(
            mask.unsqueeze(1).unsqueeze(2), float("-inf")).type_as(s).view(B * heads, L, L)
    if log_penalty:
        idx = torch.arange(L)
        dist = (idx.unsqueeze(1) - idx.unsqueeze(0)).abs().float()
        # max(0, log d): log 0 = -inf clamps to 0
        s = s - torch.clamp(torch.log(dist), min=0.0)
    p = F.softmax(s.float(), dim=-1).type_as(s)
    o = torch.bmm(p, v).transpose(0, 1).contiguous().view(L, B, D)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


# --------------------------------------------------------------------------- a6
def encoder_layer(sd: Dict[str, Tensor], prefix: str, x: Tensor, mask: Optional[Tensor],
                  heads: int, log_penalty: bool, ln_eps: float = 1e-5,
                  relu_masks: Optional[dict] = None) -> Tensor:
    """fairseq/modules/transformer_layer.py:87-139 with normalize_before=True,
    eval mode (all dropouts off)."""
    D = x.shape[-1]
    r = x
    x = F.layer_norm(x, (D,), sd[prefix + "self_attn_layer_norm.weight"],
                     sd[prefix + "self_attn_layer_norm.bias"], ln_eps)
    x = r + self_attention(sd, prefix + "self_attn.", x, mask, heads, log_penalty)
    r = x
    x = F.layer_norm(x, (D,), sd[prefix + "final_layer_norm.weight"],
                     sd[prefix + "final_layer_norm.bias"], ln_eps)
    x = _relu(F.linear(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]), relu_masks, prefix + "fc1")
    x = F.linear(x, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])
    return r + x


# ---------------------------------------------------------------------- a10/a11
def ctc_segments(logits: Tensor, lengths: Tensor) -> List[List[Tuple[int, int]]]:
    """ST/models/conv_transformer.py:282-285 -- per utterance, argmax over the
    vocabulary of softmax(logits) for the first len_b frames, then run-length
    groups [(label, run)].  logits is T x B x V."""
    prob = F.softmax(logits, dim=-1).transpose(0, 1)
    out = []
    for b in range(prob.shape[0]):
        pred = prob[b][: int(lengths[b])].argmax(-1).tolist()
        out.append([(lab, len(list(g))) for lab, g in groupby(pred)])
    return out


def ctc_weights(prob: Tensor, segments, strategy: str, dtype) -> Tensor:
    """ST/models/conv_transformer.py:385-426 -- dense B x T x T'' weight matrix.
    avg: 1/run.  weighted: prob[t,label]/sum.  softmax: softmax over the run of
    the PROBABILITIES prob[t,label] (then /sum, which is 1)."""
    B, T, _ = prob.shape
    new_max = max(len(s) for s in segments)
    W = torch.zeros((B, T, new_max), dtype=dtype)
    for b, segs in enumerate(segments):
        a = 0
        for s, (lab, run) in enumerate(segs):
            e = a + run
            if strategy == "avg":
                W[b, a:e, s] = 1.0 / run
            else:
                w = prob[b, a:e, lab]
                if strategy == "softmax":
                    w = F.softmax(w, dim=0)
                elif strategy != "weighted":
                    raise ValueError(strategy)
                W[b, a:e, s] = w / w.sum()
            a = e
    return W


def ctc_compress(x: Tensor, logits: Tensor, lengths: Tensor, strategy: str):
    """ST/models/conv_transformer.py:278-291 given the CTC logits (``x_ctc``).
    Returns (compressed T'' x B x D, new_lengths B int64, segments)."""
    segments = ctc_segments(logits, lengths)
    new_lengths = [len(s) for s in segments]
    prob = F.softmax(logits, dim=-1).transpose(0, 1)
    W = ctc_weights(prob, segments, strategy, x.dtype)
    out = x.permute(1, 2, 0).bmm(W).permute(2, 0, 1)
    return out, lengths.new_tensor(new_lengths), segments


# --------------------------------------------------------------------------- a12
def encoder_forward(sd: Dict[str, Tensor], cfg: dict, src_tokens: Tensor, src_lengths: Tensor,
                    return_all_hiddens: bool = False, ctc_logits_hook=None,
                    relu_masks: Optional[dict] = None) -> dict:
    """ST/models/conv_transformer.py:195-276 in eval mode.

    ``cfg`` keys: embed_dim, heads, layers, distance_penalty ('log' or None),
    ctc_layer (0 = no compression), ctc_strategy.  ``ctc_logits_hook(logits)``
    plays the role of an ``nn.Module`` forward hook on ``ctc_fc`` (used by tests
    and the bench to inject run-structured logits).
    Returns a dict with the fields of ``CTCAwareEncoderOut`` (:28-32).
    """
    heads = cfg["heads"]
    log_pen = cfg.get("distance_penalty", "log") == "log"
    x, lengths = conv_subsample(sd, src_tokens, src_lengths, relu_masks=relu_masks)
    x = flatten_fc3(sd, x, relu_masks)
    D = x.shape[-1]
    x = x + positional_embedding(lengths, D).to(x.dtype).transpose(0, 1)
    mask = create_mask(lengths)
    states = [] if return_all_hiddens else None
    ctc_out, ctc_mask, segments = None, None, None
    for l in range(cfg["layers"]):
        x = encoder_layer(sd, "layers.%d." % l, x, mask, heads, log_pen, relu_masks=relu_masks)
        if cfg.get("ctc_layer", 0) == l + 1:
            ctc_mask = mask
            ctc_out = F.linear(x, sd["ctc_fc.weight"], sd["ctc_fc.bias"])
            if ctc_logits_hook is not None:
                ctc_out = ctc_logits_hook(ctc_out)
            x, lengths, segments = ctc_compress(x, ctc_out, lengths, cfg.get("ctc_strategy", "avg"))
            mask = create_mask(lengths)
        if return_all_hiddens:
            states.append(x)
    if "layer_norm.weight" in sd:
        x = F.layer_norm(x, (D,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)
        if return_all_hiddens:
            states[-1] = x
    return dict(encoder_out=x, encoder_padding_mask=mask, encoder_embedding=None,
                encoder_states=states, src_tokens=src_tokens, src_lengths=lengths,
                ctc_out=ctc_out, ctc_padding_mask=ctc_mask, segments=segments)


# ------------------------------------------------------------------ a14 (init)
def init_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init weights of the reference architecture without fairseq:
    Conv2D N(0, sqrt(4(1-p)/(k*Cin))) (ST/models/conv_transformer.py:348-354),
    Linear/LocalAttention xavier_uniform (:371-375, local_attention.py:42-47),
    BatchNorm running stats randomised so BN is not the identity.
    Distribution-level restatement: NOT the same RNG stream as the reference ctor
    (parity tests take weights from the live reference via ``make_golden``)."""
    g = torch.Generator().manual_seed(seed)
    D, Dff, C, Fdim = cfg["embed_dim"], cfg["ffn_dim"], cfg.get("conv_channels", 64), cfg["feat_dim"]
    V, p = cfg["vocab"], cfg.get("dropout", 0.1)

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    def small(n):
        return 0.05 * torch.randn(n, generator=g)

    sd = {}
    cin = 1
    for i in range(2):
        std = math.sqrt(4 * (1.0 - p) / (3 * cin))
        sd["convolutions.%d.weight" % i] = torch.randn(C, cin, 3, 3, generator=g) * std
        sd["convolutions.%d.bias" % i] = small(C)
        sd["bn.%d.weight" % i] = 1.0 + 0.1 * torch.randn(C, generator=g)
        sd["bn.%d.bias" % i] = 0.1 * torch.randn(C, generator=g)
        sd["bn.%d.running_mean" % i] = 0.1 * torch.randn(C, generator=g)
        sd["bn.%d.running_var" % i] = torch.rand(C, generator=g) + 0.5
        sd["bn.%d.num_batches_tracked" % i] = torch.tensor(0)
        cin = C
    f2 = math.ceil(math.ceil(Fdim / 2) / 2)
    sd["fc3.weight"] = xavier(D, C * f2)
    sd["fc3.bias"] = small(D)
    sd["embed_positions.embeddings._float_tensor"] = torch.zeros(1)
    for l in range(cfg["layers"]):
        pre = "layers.%d." % l
        if cfg.get("distance_penalty", "log") == "log":
            sd[pre + "self_attn.in_proj_weight"] = xavier(3 * D, D)
            sd[pre + "self_attn.in_proj_bias"] = small(3 * D)
        else:
            for n in "qkv":
                sd[pre + "self_attn.%s_proj.weight" % n] = xavier(D, D)
                sd[pre + "self_attn.%s_proj.bias" % n] = small(D)
        sd[pre + "self_attn.out_proj.weight"] = xavier(D, D)
        sd[pre + "self_attn.out_proj.bias"] = small(D)
        for ln in ("self_attn_layer_norm", "final_layer_norm"):
            sd[pre + ln + ".weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            sd[pre + ln + ".bias"] = small(D)
        sd[pre + "fc1.weight"] = xavier(Dff, D)
        sd[pre + "fc1.bias"] = small(Dff)
        sd[pre + "fc2.weight"] = xavier(D, Dff)
        sd[pre + "fc2.bias"] = small(D)
    sd["layer_norm.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
    sd["layer_norm.bias"] = small(D)
    if cfg.get("ctc_layer", 0) > 0:
        a = 1.0 / math.sqrt(D)  # nn.Linear default init bound (conv_transformer.py:190)
        sd["ctc_fc.weight"] = (torch.rand(V, D, generator=g) * 2 - 1) * a
        sd["ctc_fc.bias"] = (torch.rand(V, generator=g) * 2 - 1) * a
    return sd


# --------------------------------------------------- synthetic inputs
def synthetic_batch(lengths: List[int], feat_dim: int, seed: int = 1234):
    """``src_tokens ~ N(0,1)`` B x T x F, zeroed past each length (collater
    semantics, ST/data/collaters.py:51-56), lengths as given (sorted descending)."""
    g = torch.Generator().manual_seed(seed)
    B, T = len(lengths), max(lengths)
    x = torch.randn(B, T, feat_dim, generator=g)
    for b, n in enumerate(lengths):
        x[b, n:] = 0
    return x, torch.tensor(lengths, dtype=torch.long)


def synthetic_ctc_bump(L: int, B: int, vocab: int, seed: int = 7, mean_run: float = 3.0,
                       blank_prob: float = 0.5):
    """Run-structured label plan: per utterance, labels drawn as
    runs with geometric length (mean ``mean_run``); a run is ``<ctc_blank>``
    (= vocab-1) with probability ``blank_prob``.  Returns labels L x B int64."""
    g = torch.Generator().manual_seed(seed)
    labels = torch.empty(L, B, dtype=torch.long)
    p = 1.0 / mean_run
    for b in range(B):
        t = 0
        prev = -1
        while t < L:
            u = float(torch.rand((), generator=g))
            run = 1 + int(math.log(max(u, 1e-12)) / math.log(1.0 - p))
            if float(torch.rand((), generator=g)) < blank_prob and prev != vocab - 1:
                lab = vocab - 1
            else:
                lab = int(torch.randint(4, vocab - 1, (), generator=g))
                while lab == prev:
                    lab = int(torch.randint(4, vocab - 1, (), generator=g))
            labels[t:t + run, b] = lab
            prev = lab
            t += run
    return labels


def bump_hook(labels: Tensor, margin: float):
    """Forward-hook body: logits[t,b,labels[t,b]] += margin (same on both sides)."""
    def hook(logits: Tensor) -> Tensor:
        L, B, _ = logits.shape
        lab = labels[:L, :B].to(logits.device)
        out = logits.clone()
        out.scatter_add_(2, lab.unsqueeze(-1),
                         torch.full((L, B, 1), margin, dtype=logits.dtype, device=logits.device))
        return out
    return hook
