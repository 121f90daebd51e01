"""Host-side mirror of the reference's CTC criterion arithmetic over the device path.

Reference: ``examples/speech_recognition/criterions/CTC_loss.py`` -- ``compute_ctc_uer`` (:31-74) and
the ``F.ctc_loss`` call of ``CTCCriterion.forward`` (:128-151), reached from ``ctc_multi_loss.py:27-41``
with the encoder's ``ctc_out`` (T x B x V) and ``ctc_padding_mask``.

The reference pulls every utterance's arg-max path to the host (``.tolist()``), collapses it with
``itertools.groupby`` and runs a Python O(P*U) alignment per utterance, every training and validation
step.  Here the frame arg-max, the log-sum-exp, the collapse, the alignment and the alpha recursion
all run on the GPU; the only host transfer is the final read of 3 scalars.

No CPU fallback: CPU tensors raise.
"""
import torch

from . import ops


def _time_major_rows(x):
    """[T, B, V] logits/log-probs -> ([T*B, V] row view with row t*B+b, T, B, V) without copying
    when the storage is time-major (also when `x` is the criterion's N x T x D transposed view)."""
    if x.dim() != 3:
        raise ValueError("fbkst_b200.criterion: expected a 3-D tensor")
    if not x.is_cuda:
        raise ValueError("fbkst_b200.criterion: expected a CUDA tensor (there is no CPU fallback)")
    if x.dtype not in (torch.bfloat16, torch.float32):
        x = x.float()
    T, B, V = x.shape
    if x.stride(2) != 1 or x.stride(0) != B * x.stride(1):
        x = x.contiguous()
    return x.as_strided((T * B, V), (x.stride(1), 1), x.storage_offset()), T, B, V


def _i32(t, device):
    return torch.as_tensor(t, device=device).to(torch.int32).contiguous()


def ctc_uer_device(ctc_out, targets, input_lengths, target_lengths, blank_idx):
    """Device tensors only, no synchronisation.  ctc_out T x B x V (logits or log-probs: the arg-max
    is the same).  -> (errors [B] int32, totals [2] int64 = (batch_errors, batch_total))."""
    rows, T, B, V = _time_major_rows(ctc_out)
    dev = rows.device
    il, tl = _i32(input_lengths, dev), _i32(target_lengths, dev)
    labels, _ = ops.ctc_argmax(rows, il, T, B, V, want_prob=False)
    tg = torch.as_tensor(targets, device=dev).to(torch.int64)
    errors, _, totals = ops.ctc_uer(labels, il, tg, tl, blank_idx, T, B)
    return errors, totals


def compute_ctc_uer(logprobs, targets, input_lengths, target_lengths, blank_idx):
    """Same signature and return value as the reference's ``compute_ctc_uer`` (CTC_loss.py:31-74):
    logprobs N x T1 x D (the criterion passes the transposed view of the T x N x D tensor, which is
    consumed without a copy), targets N x T2, lengths per sample.  Returns (batch_errors, batch_total)
    as Python numbers (one device->host read of 16 bytes)."""
    _, totals = ctc_uer_device(logprobs.transpose(0, 1), targets, input_lengths, target_lengths, blank_idx)
    e, n = totals.tolist()
    return float(e), float(n)


class CtcLossFn(torch.autograd.Function):
    """sum_b nll_b of F.ctc_loss(log_softmax(logits), ..., reduction="sum", zero_infinity=True), differentiable
    w.r.t. the logits on the device kernels (alpha pass forward; dense softmax term + alpha/beta occupancies
    backward).  The log-probabilities are never materialised (the reference's autograd keeps a T x B x V
    log_softmax output alive for the backward)."""

    @staticmethod
    def forward(ctx, ctc_out, in_lengths, targets, target_lengths, blank_idx):
        rows, T, B, V = _time_major_rows(ctc_out)
        labels, lse, _ = ops.ctc_argmax_lse(rows, in_lengths, T, B, V)
        nll, loss = ops.ctc_loss_fwd(rows, lse, in_lengths, targets, target_lengths, blank_idx, T, B, V)
        ctx.save_for_backward(rows, lse, in_lengths, targets, target_lengths)
        ctx.dims = (T, B, V, blank_idx, tuple(ctc_out.shape))
        ctx.mark_non_differentiable(labels, nll)
        return loss.reshape(()), labels, nll

    @staticmethod
    def backward(ctx, grad_loss, _gl, _gn):
        rows, lse, il, tg, tl = ctx.saved_tensors
        T, B, V, blank, shape = ctx.dims
        dz = ops.ctc_loss_bwd(rows, lse, il, tg, tl, blank, grad_loss, T, B, V)  # [T*B, V], pitch ceil8(V)
        return dz.view(T, B, V), None, None, None, None


class CtcProjLossFn(torch.autograd.Function):
    """``CtcLossFn`` + the backward of the CTC projection (``ctc_fc``: conv_transformer.py:278-280) in ONE node.

    Why: autograd runs nodes in reverse creation order, so the loss nodes run first, then the whole decoder
    backward (bound by the host's launch rate, GPU nearly idle), and only then the encoder's node.  With the
    projection's backward here -- d logits, d W_ctc = d logits^T x, d b_ctc, d x = d logits W_ctc, the largest
    GEMMs of the step (V = 8005) -- that work executes on the GPU WHILE the host walks the decoder's backward
    instead of on the GPU-bound stretch after it.  ``x`` is the tap the encoder's training forward hands out
    (``ctc_out._fbkst_tap``); ``logits`` are the values (already computed by the encoder, hooks applied)."""

    @staticmethod
    def forward(ctx, logits, x, weight, bias, enc, in_lengths, targets, target_lengths, blank_idx):
        rows, T, B, V = _time_major_rows(logits)
        labels, lse, _ = ops.ctc_argmax_lse(rows, in_lengths, T, B, V)
        nll, loss = ops.ctc_loss_fwd(rows, lse, in_lengths, targets, target_lengths, blank_idx, T, B, V)
        ctx.save_for_backward(rows, lse, in_lengths, targets, target_lengths, x)
        ctx.dims = (T, B, V, blank_idx)
        ctx.enc = enc
        ctx.mark_non_differentiable(labels, nll)
        return loss.reshape(()), labels, nll

    @staticmethod
    def backward(ctx, grad_loss, _gl, _gn):
        from .train import prepare_train_weights, pretranspose
        rows, lse, il, tg, tl, x = ctx.saved_tensors
        T, B, V, blank = ctx.dims
        enc = ctx.enc
        # this node runs FIRST in the backward: enqueue the encoder's activation transposes now, so that they
        # execute while the host walks the decoder's backward
        pretranspose(getattr(enc, "_train_pending", None))
        enc._train_pending = None
        W = prepare_train_weights(enc)
        D = x.shape[-1]
        M, Vp = T * B, (V + 7) // 8 * 8
        dz = ops.ctc_loss_bwd(rows, lse, il, tg, tl, blank, grad_loss, T, B, V)  # [M, V] fp32, pitch ceil8(V)
        gb, gT, dbias = ops.grad_prep(dz, n_pad=Vp)
        xb = ops.cast_bf16(x.reshape(M, D))
        dW = ops.linear_wgrad(gT, ops.transpose_bf16(xb))
        wcT = W["wcT"]
        wcT_full = wcT if Vp == V else torch.as_strided(wcT, (D, Vp), (wcT.stride(0), 1))
        dx = ops.linear(gb, wcT_full, None, out_dtype=torch.float32)
        w, b = enc.ctc_fc.weight, enc.ctc_fc.bias
        return (None, dx.view(x.shape), dW if dW.dtype == w.dtype else dW.to(w.dtype),
                dbias if dbias.dtype == b.dtype else dbias.to(b.dtype), None, None, None, None, None)


def ctc_loss_and_uer(ctc_out, ctc_padding_mask, targets, target_lengths, blank_idx, pad_idx=None):
    """What ``CTCCriterion.forward`` computes from the encoder output (CTC_loss.py:118-154):

        lprobs = log_softmax(ctc_out.float());  input_lengths from the padding mask
        loss   = F.ctc_loss(lprobs, targets, input_lengths, target_lengths, blank, "sum", zero_infinity)
        errors, total = compute_ctc_uer(lprobs, targets, input_lengths, target_lengths, blank)

    ctc_out T x B x V (bf16 or fp32 logits, ``encoder_out.ctc_out``), ctc_padding_mask B x T bool
    (True = padding) or None or a [B] tensor of lengths.  One pass over the logits produces the frame
    arg-max and the log-sum-exp; the log-probabilities are never materialised.
    Returns device tensors (loss [1] fp32, nll [B] fp32, errors [B] int32, totals [2] int64)."""
    rows, T, B, V = _time_major_rows(ctc_out)
    dev = rows.device
    if ctc_padding_mask is None:
        il = torch.full((B,), T, dtype=torch.int32, device=dev)
    elif ctc_padding_mask.dim() == 1:
        il = _i32(ctc_padding_mask, dev)
    else:  # data_utils.encoder_padding_mask_to_lengths: T - number of padded positions
        m = ctc_padding_mask if ctc_padding_mask.shape[0] == B else ctc_padding_mask.t()
        il = (T - m.sum(dim=1)).to(torch.int32)
    tl = _i32(target_lengths, dev)
    tg = torch.as_tensor(targets, device=dev).to(torch.int64)
    labels, lse, _ = ops.ctc_argmax_lse(rows, il, T, B, V)
    nll, loss = ops.ctc_loss_fwd(rows, lse, il, tg, tl, blank_idx, T, B, V)
    errors, _, totals = ops.ctc_uer(labels, il, tg, tl, blank_idx, T, B)
    return loss, nll, errors, totals


def ctc_loss_train(ctc_out, ctc_padding_mask, targets, target_lengths, blank_idx, mask_time_first=True):
    """Differentiable CTC loss + UER counts for the training criterion.  ctc_out T x B x V logits (attached to
    the autograd graph), ctc_padding_mask bool (True = padding), T x B if ``mask_time_first`` (what
    CTCEncoderWrapperModel hands over, ctc_multi_loss.py:41) else B x T, or None.
    -> (loss 0-dim tensor with grad_fn, errors total (device int64 [2]), input_lengths int32 [B])."""
    T, B, V = ctc_out.shape
    dev = ctc_out.device
    if ctc_padding_mask is None:
        il = torch.full((B,), T, dtype=torch.int32, device=dev)
    else:
        il = (T - ctc_padding_mask.sum(dim=0 if mask_time_first else 1)).to(torch.int32)
    tl = _i32(target_lengths, dev)
    tg = torch.as_tensor(targets, device=dev).to(torch.int64).contiguous()
    tap = getattr(ctc_out, "_fbkst_tap", None)
    if tap is not None and tap[0].requires_grad and tap[1].ctc_fc.weight.requires_grad:
        x, enc = tap  # the encoder's own training forward produced these logits: projection backward done here
        loss, labels, _ = CtcProjLossFn.apply(ctc_out.detach(), x, enc.ctc_fc.weight, enc.ctc_fc.bias, enc, il, tg,
                                              tl, int(blank_idx))
    else:
        loss, labels, _ = CtcLossFn.apply(ctc_out, il, tg, tl, int(blank_idx))
    _, _, totals = ops.ctc_uer(labels, il, tg, tl, blank_idx, T, B)
    return loss, totals, il
