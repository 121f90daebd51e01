"""CPU restatement of the decoder's encoder-decoder attention on its incremental (static_kv) path and
of the beam bookkeeping around it -- TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by
``tests/``, never by the product.

Pinned by ``tests/golden/xattn.pt``: outputs of the LIVE reference ``MultiheadAttention`` /
``reorder_incremental_state`` / ``reorder_encoder_out`` (generator: ``oracle/make_golden_xattn.py``).

Reference lines followed (``/root/reference``):
  fairseq/modules/multihead_attention.py:194-209   q/k/v projections, q *= head_dim**-0.5
  :229-246                                         (T, B, D) -> (B*H, T, hd) head split
  :248-281                                         static_kv cache: k, v, key-padding mask saved per
                                                   hypothesis row as (bsz, H, S, hd)
  :317-335                                         scores = q k^T ; padding keys -> -inf
  :340-353                                         fp32 softmax ; attn = P v ; out_proj
  :355-362                                         weights (H, bsz, tgt, S) or their mean over heads
  :407-420                                         reorder_incremental_state: index_select(0, new_order)
                                                   unless the cached batch already has that size
  examples/speech_recognition/models/conv_transformer.py:315-345   reorder_encoder_out
"""
import torch
import torch.nn.functional as F


def init_params(D, kdim, seed):
    """Parameters of one encoder-decoder attention block (names as in the reference state_dict)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"q_proj.weight": r(D, D) * D ** -0.5, "q_proj.bias": r(D) * 0.1,
            "k_proj.weight": r(D, kdim) * kdim ** -0.5, "k_proj.bias": r(D) * 0.1,
            "v_proj.weight": r(D, kdim) * kdim ** -0.5, "v_proj.bias": r(D) * 0.1,
            "out_proj.weight": r(D, D) * D ** -0.5, "out_proj.bias": r(D) * 0.1}


def _split_heads(x, H):
    """(T, B, D) -> (B*H, T, hd)   multihead_attention.py:229-246"""
    T, B, D = x.shape
    return x.contiguous().view(T, B * H, D // H).transpose(0, 1)


def cross_attention(P, H, query, key, key_padding_mask, state=None, need_weights=True,
                    need_head_weights=False):
    """One call of the block.  ``state`` (a dict, mutated) is the layer's incremental buffer: when it
    already holds ``prev_key`` the cached, per-hypothesis K/V are used and ``key`` is ignored
    (static_kv, :181-186, :255-262).  Returns (attn (tgt, bsz, D), weights or None)."""
    tgt, bsz, D = query.shape
    hd = D // H
    q = F.linear(query, P["q_proj.weight"], P["q_proj.bias"]) * hd ** -0.5
    q = _split_heads(q, H)
    if state is not None and "prev_key" in state:
        k = state["prev_key"].view(bsz * H, -1, hd)
        v = state["prev_value"].view(bsz * H, -1, hd)
        key_padding_mask = state.get("prev_key_padding_mask")  # :377-378
    else:
        k = _split_heads(F.linear(key, P["k_proj.weight"], P["k_proj.bias"]), H)
        v = _split_heads(F.linear(key, P["v_proj.weight"], P["v_proj.bias"]), H)
    if state is not None:
        state["prev_key"] = k.view(bsz, H, -1, hd)
        state["prev_value"] = v.view(bsz, H, -1, hd)
        state["prev_key_padding_mask"] = key_padding_mask
    S = k.size(1)
    w = torch.bmm(q, k.transpose(1, 2))  # (bsz*H, tgt, S)
    if key_padding_mask is not None:
        w = w.view(bsz, H, tgt, S).masked_fill(
            key_padding_mask.view(bsz, 1, 1, S).to(torch.bool), float("-inf")).view(bsz * H, tgt, S)
    p = torch.softmax(w.float(), dim=-1)
    a = torch.bmm(p.type_as(w), v)  # (bsz*H, tgt, hd)
    a = a.transpose(0, 1).contiguous().view(tgt, bsz, D)
    a = F.linear(a, P["out_proj.weight"], P["out_proj.bias"])
    weights = None
    if need_weights:
        weights = p.view(bsz, H, tgt, S).transpose(1, 0)
        if not need_head_weights:
            weights = weights.mean(dim=0)
    return a, weights


def reorder_state(state, new_order):
    """multihead_attention.py:407-420 for an encoder-decoder block: every cached tensor is gathered
    along the hypothesis dimension, except that NOTHING is touched once the first cached tensor
    already has ``new_order``'s batch size (the reference ``break``s out of the loop)."""
    for k in list(state.keys()):
        t = state[k]
        if t is not None:
            if t.size(0) == new_order.size(0):
                break
            state[k] = t.index_select(0, new_order)
    return state


def reorder_encoder_out(encoder_out, encoder_padding_mask, new_order):
    """conv_transformer.py:329-338: T x B x C gathered on dim 1, the B x T mask on dim 0."""
    eo = encoder_out.index_select(1, new_order)
    m = None if encoder_padding_mask is None else encoder_padding_mask.index_select(0, new_order)
    return eo, m
