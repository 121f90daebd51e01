"""GPU parity of the training-side kernels (scope row T), through the C ABI, against torch.autograd of
the same fp32 formula on the same (bf16-rounded) inputs.

Tolerances: fp32 streaming kernels 1e-5; bf16-output kernels 1e-2 of the tensor's max magnitude (one bf16
rounding of the result); tensor-core gradients (bf16 operands, fp32 accumulation) 2e-2 (north_star)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402


def dev():
    return torch.device("cuda:0")


def bf(x):
    return x.to(torch.bfloat16)


# ----------------------------------------------------------------------- dropout + residual + LayerNorm
@pytest.mark.parametrize("M,D", [(100, 128), (777, 512), (64, 1024)])
def test_dropout_add_ln_no_dropout(M, D):
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(M + D)
    y, r = torch.randn(M, D, generator=g).to(dev()), torch.randn(M, D, generator=g).to(dev())
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).to(dev()), (0.1 * torch.randn(D, generator=g)).to(dev())
    x1, ln = ops.dropout_add_ln(y, r, gamma, beta)
    assert torch.equal(x1, y + r)
    ref = torch.nn.functional.layer_norm(y + r, (D,), gamma, beta, 1e-5)
    assert rel_err(ln.float(), ref) < 1e-2
    x1b, ln_none = ops.dropout_add_ln(y, None)
    assert ln_none is None and torch.equal(x1b, y)


def test_dropout_mask_forward_backward_agree():
    """The backward regenerates the forward's mask: grad_prep with the same (seed, site) on a tensor of
    ones reproduces exactly the keep pattern dropout_add_ln applied; the keep rate is 1 - p."""
    from fbkst_b200 import ops
    M, D, p = 513, 512, 0.3
    y = torch.ones(M, D, device=dev())
    x1, _ = ops.dropout_add_ln(y, None, p=p, seed=1234567891011, site=7)
    kept = x1 != 0
    assert abs(kept.float().mean().item() - (1 - p)) < 5e-3
    assert torch.allclose(x1[kept], torch.full_like(x1[kept], 1 / (1 - p)))
    gb, gT, cs = ops.grad_prep(torch.ones(M, D, device=dev()), p=p, seed=1234567891011, site=7)
    assert torch.equal(gb.float() != 0, kept)
    assert torch.equal(gT.t().float() != 0, kept)
    x2, _ = ops.dropout_add_ln(y, None, p=p, seed=1234567891011, site=8)  # another site: another mask
    assert not torch.equal(x2 != 0, kept)
    z = torch.ones(M, D, device=dev(), dtype=torch.bfloat16)
    ops.dropout_(z, p, 1234567891011, 7)  # same counter layout for the in-place kernel
    assert torch.equal(z != 0, kept)


@pytest.mark.parametrize("M,D", [(100, 128), (3000, 512), (257, 1024)])
def test_ln_bwd(M, D):
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(M * 3 + D)
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).to(dev()).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).to(dev()).requires_grad_(True)
    beta = (0.1 * torch.randn(D, generator=g)).to(dev()).requires_grad_(True)
    dy = torch.randn(M, D, generator=g).to(dev())
    res = torch.randn(M, D, generator=g).to(dev())
    torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-5).backward(dy)
    dx, dg, db = ops.ln_bwd(dy, x.detach(), gamma.detach(), dx=res.clone())
    assert rel_err(dx, x.grad + res) < 1e-5
    assert rel_err(dg, gamma.grad) < 1e-4 and rel_err(db, beta.grad) < 1e-4
    dx2, _, _ = ops.ln_bwd(dy, x.detach(), gamma.detach())
    assert rel_err(dx2, x.grad) < 1e-5


# ----------------------------------------------------------------------- grad_prep (mask, transpose, sums)
@pytest.mark.parametrize("M,N,f32", [(64, 64, True), (1000, 512, True), (333, 2048, False), (130, 8005, True)])
def test_grad_prep(M, N, f32):
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, N, generator=g).to(dev())
    if not f32:
        x = bf(x)
    act = bf(torch.relu(torch.randn(M, N, generator=g))).to(dev())
    n_pad = (N + 7) // 8 * 8
    gb, gT, cs = ops.grad_prep(x, act=act, act_scale=1.25, n_pad=n_pad)
    ref = bf(torch.where(act > 0, x.float() * 1.25, torch.zeros_like(x.float())))
    assert gb.shape == (M, n_pad) and torch.equal(gb[:, :N], ref) and (gb[:, N:] == 0).all()
    assert gT.shape == (N, M) and torch.equal(gT, ref.t())
    assert rel_err(cs, ref.float().sum(0)) < 1e-5
    t = ops.transpose_bf16(x)
    assert torch.equal(t, bf(x).t())


def test_grad_prep_row_remap():
    """fc3's backward: the incoming gradient is time-major (t*B+b), the GEMM operands are in the conv
    layout's (b*L+t) row order."""
    from fbkst_b200 import ops
    L, B, N = 37, 5, 256
    g = torch.randn(L * B, N, device=dev())
    act = bf(torch.relu(torch.randn(B * L, N, device=dev())))
    gb, gT, cs = ops.grad_prep(g, act=act, remap=(L, B))
    ref = bf(torch.where(act > 0, g.view(L, B, N).transpose(0, 1).reshape(B * L, N), torch.zeros_like(g)))
    assert torch.equal(gb, ref) and torch.equal(gT, ref.t())


# ----------------------------------------------------------------------- wgrad (split-K tcgen05 GEMM)
@pytest.mark.parametrize("n_out,k_in,tokens", [(512, 512, 1000), (1536, 512, 24000), (2048, 512, 4099),
                                                (8005, 512, 3000), (64, 576, 30000), (512, 640, 777)])
def test_linear_wgrad(n_out, k_in, tokens):
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(n_out + tokens)
    dy = bf(torch.randn(tokens, n_out, generator=g)).to(dev())
    x = bf(torch.randn(tokens, k_in, generator=g)).to(dev())
    dW = ops.linear_wgrad(ops.transpose_bf16(dy), ops.transpose_bf16(x))
    ref = dy.float().t() @ x.float()
    assert dW.shape == (n_out, k_in)
    assert rel_err(dW, ref) < 1e-4
    dW2 = ops.linear_wgrad(ops.transpose_bf16(dy), ops.transpose_bf16(x))
    assert torch.equal(dW, dW2)  # fixed summation order


@pytest.mark.parametrize("n_out,k_in,tokens,x_f16", [(512, 512, 1000, False), (1536, 512, 24000, False),
                                                      (2048, 512, 4099, False), (8005, 512, 3000, False),
                                                      (64, 576, 30000, False), (512, 640, 777, False),
                                                      (512, 2048, 24000, False), (200, 72, 130, False)])
def test_linear_wgrad_nt(n_out, k_in, tokens, x_f16):
    """dW = g^T x with both operands read MN-major from their natural layouts (fbkst_linear_wgrad_nt), pitch-padded
    operands included; must equal the transposed-copy route bit for bit (same tiles, same split-K order)."""
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(n_out + tokens)
    ldg, ldx = (n_out + 7) // 8 * 8, (k_in + 7) // 8 * 8
    dy = torch.zeros(tokens, ldg, dtype=torch.bfloat16, device=dev())[:, :n_out]
    dy.copy_(bf(torch.randn(tokens, n_out, generator=g)))
    x = torch.zeros(tokens, ldx, dtype=torch.float16 if x_f16 else torch.bfloat16, device=dev())[:, :k_in]
    x.copy_(torch.randn(tokens, k_in, generator=g))
    dW = ops.linear_wgrad_nt(dy, x)
    ref = dy.float().t() @ x.float()
    assert dW.shape == (n_out, k_in)
    assert rel_err(dW, ref) < 1e-4
    if not x_f16:
        dW2 = ops.linear_wgrad(ops.transpose_bf16(dy), ops.transpose_bf16(x))
        assert torch.equal(dW, dW2)


# ----------------------------------------------------------------------- attention (training)
def attn_train_ref(qkv, lengths, L, B, H, log_penalty, keep=None):
    """local_attention.py:98-139 in fp32, differentiable; keep [B*H, L, L] = dropout keep-scale."""
    D = H * 64
    q, k, v = qkv.view(L, B, 3, H, 64).permute(2, 1, 3, 0, 4)  # [3] B H L 64
    s = (q * 0.125) @ k.transpose(-1, -2)
    key_pad = torch.arange(L, device=qkv.device)[None, :] >= lengths[:, None]
    s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    if log_penalty:
        i = torch.arange(L, device=qkv.device)
        s = s - torch.clamp(torch.log((i[:, None] - i[None, :]).abs().float()), min=0)
    p = torch.softmax(s, -1)
    if keep is not None:
        p = p * keep.view(B, H, L, L)
    o = p @ v
    return o.permute(2, 0, 1, 3).reshape(L * B, D)


def valid_rows(lengths, L, B):
    return (torch.arange(L, device=lengths.device)[:, None] < lengths[None, :]).reshape(L * B)


@pytest.mark.parametrize("L,B,H,lens,pen", [
    (64, 1, 1, [64], False), (128, 2, 2, [128, 128], True), (100, 3, 2, [100, 64, 5], True),
    (375, 2, 8, [375, 201], True), (300, 2, 4, [300, 129], False), (700, 1, 2, [700], True)])
def test_attention_train_fwd_bwd(L, B, H, lens, pen):
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(L + B + H)
    D = H * 64
    qkv = bf(torch.randn(L * B, 3 * D, generator=g)).to(dev())
    lengths = torch.tensor(lens, dtype=torch.int32, device=dev())
    out, lse = ops.attention_train_fwd(qkv, lengths, L, B, H, pen)
    x = qkv.float().requires_grad_(True)
    ref = attn_train_ref(x, lengths, L, B, H, pen)
    vr = valid_rows(lengths, L, B)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.float()[vr], ref.detach()[vr]) < 1e-2
    # the inference kernel computes the same thing from pre-scaled q
    qs = qkv.clone()
    qs[:, :D] = bf(qkv[:, :D].float() * 0.125)
    inf = ops.attention(qs, lengths, L, B, H, pen)
    assert rel_err(out.float()[vr], inf.float()[vr]) < 1e-2
    # backward: padded query rows carry no gradient (the loss never reads them)
    dout = bf(torch.randn(L * B, D, generator=g)).to(dev())
    dout[~vr] = 0
    ref.backward(dout.float())
    dqkv = ops.attention_train_bwd(qkv, out, dout, lse, lengths, L, B, H, pen)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        e = rel_err(dqkv[:, sl].float(), x.grad[:, sl])
        assert e < 2e-2, (name, e)
    assert (dqkv.float()[~vr] == 0).all()  # padded rows: exact zeros in dq, dk, dv


def test_attention_train_dropout_consistency():
    """Attention dropout: the mask is a pure function of (seed, site, head-batch, q, k).  Extract it with
    q = 0 (uniform probabilities) and v = one-hot keys, then check forward and backward against the fp32
    formula using that mask."""
    from fbkst_b200 import ops
    L, B, H, p, seed, site = 64, 2, 2, 0.25, 987654321, 3
    D = H * 64
    lengths = torch.tensor([64, 64], dtype=torch.int32, device=dev())
    probe = torch.zeros(L, B, 3, H, 64, device=dev())
    probe[:, :, 2] = torch.eye(64, device=dev())[:, None, None, :]  # v[t] = e_t for every (b, h)
    o, _ = ops.attention_train_fwd(bf(probe.view(L * B, 3 * D)), lengths, L, B, H, False, p=p, seed=seed, site=site)
    keep = (o.float().view(L, B, H, 64) * 64).permute(1, 2, 0, 3).reshape(B * H, L, L)  # (1/64) * keep -> keep
    frac = (keep == 0).float().mean().item()
    assert abs(frac - p) < 0.03
    assert torch.allclose(keep[keep != 0], torch.full_like(keep[keep != 0], 1 / (1 - p)), rtol=1e-2)
    keep = torch.where(keep != 0, torch.full_like(keep, 1 / (1 - p)), torch.zeros_like(keep))
    g = torch.Generator().manual_seed(5)
    qkv = bf(torch.randn(L * B, 3 * D, generator=g)).to(dev())
    out, lse = ops.attention_train_fwd(qkv, lengths, L, B, H, True, p=p, seed=seed, site=site)
    x = qkv.float().requires_grad_(True)
    ref = attn_train_ref(x, lengths, L, B, H, True, keep=keep)
    assert rel_err(out.float(), ref.detach()) < 1e-2
    dout = bf(torch.randn(L * B, D, generator=g)).to(dev())
    ref.backward(dout.float())
    dqkv = ops.attention_train_bwd(qkv, out, dout, lse, lengths, L, B, H, True, p=p, seed=seed, site=site)
    assert rel_err(dqkv.float(), x.grad) < 2e-2


# ----------------------------------------------------------------------- CTC compression backward
def test_ctc_compress_bwd():
    from fbkst_b200 import ops
    L, B, D, V = 50, 3, 256, 40
    g = torch.Generator().manual_seed(9)
    lengths = torch.tensor([50, 33, 7], dtype=torch.int32, device=dev())
    logits = torch.randn(L * B, V, generator=g)
    runs = torch.randint(0, V, (L // 3 + 1, B), generator=g).repeat_interleave(3, 0)[:L]
    logits.view(L, B, V).scatter_add_(2, runs.unsqueeze(-1), torch.full((L, B, 1), 20.0))
    logits = logits.to(dev())
    x = torch.randn(L * B, D, generator=g).to(dev())
    labels, prob = ops.ctc_argmax(logits, lengths, L, B, V, True)
    seg_id, seg_start, weight, new_len, max_new = ops.ctc_segment(labels, prob, lengths, "weighted", L, B)
    out = ops.ctc_compress(x, seg_id, seg_start, weight, lengths, new_len, max_new, L, B)
    dout = torch.randn(L * B, D, generator=g).to(dev())
    dx = ops.ctc_compress_bwd(dout, seg_id, weight, L, B)
    # dense restatement: out[s, b] = sum_t W[t, s] x[t, b]  (conv_transformer.py:290)
    xr = x.clone().requires_grad_(True)
    W = torch.zeros(B, L, L, device=dev())
    sid, w = seg_id.view(L, B), weight.view(L, B)
    for b in range(B):
        for t in range(int(lengths[b])):
            W[b, t, sid[t, b]] = w[t, b]
    ref = torch.einsum("bts,tbd->sbd", W, xr.view(L, B, D)).reshape(L * B, D)
    n_rows = int(max_new.item()) * B  # rows beyond the longest compressed utterance are never written
    assert rel_err(out[:n_rows], ref.detach()[:n_rows]) < 1e-5
    dout[n_rows:] = 0
    dx = ops.ctc_compress_bwd(dout, seg_id, weight, L, B)
    ref.backward(dout)
    assert rel_err(dx, xr.grad) < 1e-5


# ----------------------------------------------------------------------- BatchNorm (training) + conv pieces
@pytest.mark.parametrize("P,C", [(5000, 64), (777, 128)])
def test_bn_train_forward_backward(P, C):
    """nn.BatchNorm2d in training mode after a ReLU (conv_transformer.py:212): batch statistics over every
    pixel, running-stat update (momentum 0.1, unbiased variance), and the autograd of BN(relu(z))."""
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(P + C)
    z = (torch.randn(P, C, generator=g) * 1.5 + 0.3).half().to(dev())
    r = torch.relu(z)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev())
    beta = (0.1 * torch.randn(C, generator=g)).to(dev())
    rm, rv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    mean, rstd, sc, sh = ops.bn_batch_stats(r, gamma, beta, 1e-5, 0.1, rm, rv)
    y = ops.bn_apply(r, sc, sh)
    zr = z.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    yr = torch.nn.functional.batch_norm(torch.relu(zr).t().reshape(1, C, P, 1), rm_ref, rv_ref, gr, br, True, 0.1,
                                        1e-5).reshape(C, P).t()
    assert rel_err(y.float(), yr.detach()) < 2e-3
    assert rel_err(rm, rm_ref) < 1e-4 and rel_err(rv, rv_ref) < 1e-4
    dy = bf(torch.randn(P, C, generator=g)).to(dev())
    (yr * dy.float()).sum().backward()
    dz, dbeta, dgamma = ops.bn_relu_bwd(dy, r, gamma, mean, rstd, True)
    assert rel_err(dz.float(), zr.grad) < 1e-2
    assert rel_err(dgamma, gr.grad) < 1e-3 and rel_err(dbeta, br.grad) < 1e-3
    # dropout: the backward regenerates bn_apply's mask
    yd = ops.bn_apply(torch.ones_like(r), torch.ones(C, device=dev()), torch.zeros(C, device=dev()), 0.2, 77, 1)
    kept = yd != 0
    assert abs(kept.float().mean().item() - 0.8) < 1e-2
    dzd, _, _ = ops.bn_relu_bwd(torch.ones(P, C, device=dev(), dtype=torch.bfloat16), torch.ones_like(r),
                                torch.ones(C, device=dev()), torch.zeros(C, device=dev()), torch.ones(C, device=dev()),
                                False, 0.2, 77, 1)
    assert torch.equal(dzd != 0, kept)


@pytest.mark.parametrize("B,T1,F1,C", [(2, 31, 20, 64), (3, 50, 20, 64), (2, 23, 40, 128)])
def test_conv2_backward_pieces(B, T1, F1, C):
    """conv2's weight and input gradients (im2col^T -> split-K GEMM; GEMM -> col2im) vs autograd of F.conv2d."""
    from fbkst_b200 import ops
    g = torch.Generator().manual_seed(B * T1 + C)
    y1 = torch.randn(B, T1, F1, C, generator=g).half().to(dev())
    w2 = (torch.randn(C, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(dev())
    T2, F2 = (T1 + 1) // 2, (F1 + 1) // 2
    dz2 = bf(torch.randn(B, T2, F2, C, generator=g)).to(dev())
    xr = y1.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = bf(w2).float().requires_grad_(True)
    out = torch.nn.functional.conv2d(xr, wr, None, stride=2, padding=1)
    out.backward(dz2.float().permute(0, 3, 1, 2))
    P2 = B * T2 * F2
    colT = ops.conv2_im2col_t(y1)
    ref_col = torch.nn.functional.unfold(y1.float().permute(0, 3, 1, 2), 3, padding=1, stride=2)  # B, C*9, T2*F2
    ref_colT = ref_col.view(B, C, 9, T2 * F2).permute(2, 1, 0, 3).reshape(9 * C, P2)
    assert torch.equal(colT.float(), bf(ref_colT).float())
    dW2p = ops.linear_wgrad(ops.transpose_bf16(dz2.view(P2, C)), colT)
    dW2 = dW2p.view(C, 3, 3, C).permute(0, 3, 1, 2)
    assert rel_err(dW2, wr.grad) < 5e-3
    w2d = ops.cast_bf16(w2.permute(2, 3, 1, 0).reshape(9 * C, C).contiguous())
    dcol = ops.linear(dz2.view(P2, C), w2d)
    dy1 = ops.conv2_col2im(dcol, B, T1, F1, C)
    assert rel_err(dy1.float().permute(0, 3, 1, 2), xr.grad) < 2e-2


def test_conv1_wgrad():
    from fbkst_b200 import ops
    B, T, Fd, C = 3, 61, 40, 64
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, Fd, generator=g).to(dev())
    T1, F1 = (T + 1) // 2, (Fd + 1) // 2
    dz1 = bf(torch.randn(B, T1, F1, C, generator=g)).to(dev())
    w = torch.randn(C, 1, 3, 3, generator=g).to(dev()).requires_grad_(True)
    b = torch.zeros(C, device=dev(), requires_grad=True)
    torch.nn.functional.conv2d(x.unsqueeze(1), w, b, stride=2, padding=1).backward(dz1.float().permute(0, 3, 1, 2))
    dW1, db1 = ops.conv1_wgrad(dz1, x)
    assert rel_err(dW1.reshape(C, 1, 3, 3), w.grad) < 1e-4
    assert rel_err(db1, b.grad) < 1e-4


def test_prep_batch_many_jobs():
    """fbkst_prep_batch: > 48 (copy, transposed) jobs of mixed source types and ragged shapes in one call,
    including row-block jobs that fill one destination (the q / k / v layout) and a zero-padded transposed pitch."""
    from fbkst_b200 import ops as O
    torch.manual_seed(3)
    g = torch.Generator(device="cuda").manual_seed(4)
    jobs, checks = [], []
    shapes = [(64, 64), (130, 72), (512, 1536), (77, 8), (1, 200), (300, 1), (2048, 512)]
    for i in range(55):
        r, c = shapes[i % len(shapes)]
        dt = (torch.float32, torch.bfloat16, torch.float16)[i % 3]
        src = (torch.randn(r, c, device="cuda", generator=g) * 3).to(dt)
        cp = torch.empty(r, (c + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")[:, :c] if i % 4 != 3 else None
        tr = O.transposed_buffer(r, c, "cuda") if i % 5 != 4 or cp is None else None
        jobs.append((src, cp, tr))
        checks.append((src, cp, tr))
    # three row blocks into one destination pair
    D = 96
    blocks = [torch.randn(D, 40, device="cuda", generator=g) for _ in range(3)]
    full = torch.empty(3 * D, 40, dtype=torch.bfloat16, device="cuda")
    fullT = torch.zeros(40, 3 * D + 8, dtype=torch.bfloat16, device="cuda")
    for k, w in enumerate(blocks):
        jobs.append((w, full[k * D:(k + 1) * D], fullT[:, k * D:(k + 1) * D]))
    O.prep_batch(jobs)
    torch.cuda.synchronize()
    for src, cp, tr in checks:
        ref = src.float().to(torch.bfloat16)
        if cp is not None:
            assert torch.equal(cp, ref)
        if tr is not None:
            assert torch.equal(tr, ref.t())
    ref = torch.cat(blocks, 0).to(torch.bfloat16)
    assert torch.equal(full, ref) and torch.equal(fullT[:, :3 * D], ref.t())
    assert (fullT[:, 3 * D:] == 0).all()  # pad columns untouched


def test_deferred_reductions_match_immediate():
    """ops.deferred_reductions(): > 56 queued reductions (both kernels: few / many slices) == immediate ones."""
    from fbkst_b200 import ops as O
    g = torch.Generator(device="cuda").manual_seed(5)
    cases = []
    for i in range(70):
        G, rows, cols = [(3, 40, 24), (375, 1, 512), (10, 512, 512), (40, 2, 1024), (17, 1, 7)][i % 5]
        ldi = cols + (8 if i % 2 else 0)
        src = torch.randn(G, rows, ldi, device="cuda", generator=g)
        cases.append((src, G, rows * ldi, rows, cols, ldi))
    outs_now = []
    for src, G, gs, rows, cols, ldi in cases:
        o = torch.empty(rows, cols, device="cuda")
        O.reduce_sum(src, G, gs, rows, cols, ldi, o, cols, 0.5)
        outs_now.append(o)
    outs_q = []
    with O.deferred_reductions():
        for src, G, gs, rows, cols, ldi in cases:
            o = torch.full((rows, cols), float("nan"), device="cuda")
            O.reduce_sum(src, G, gs, rows, cols, ldi, o, cols, 0.5)
            outs_q.append(o)
    torch.cuda.synchronize()
    for (src, G, gs, rows, cols, ldi), a, b in zip(cases, outs_now, outs_q):
        assert torch.equal(a, b)  # same fixed summation order
        ref = 0.5 * src[:, :, :cols].double().sum(0)
        assert (a.double() - ref).abs().max() <= 1e-4 * max(1.0, ref.abs().max().item())
    # wgrad through the queue == wgrad without
    gT = torch.randn(256, 4096, device="cuda", generator=g).to(torch.bfloat16)
    xT = torch.randn(192, 4096, device="cuda", generator=g).to(torch.bfloat16)
    a = O.linear_wgrad(gT, xT)
    with O.deferred_reductions():
        b = O.linear_wgrad(gT, xT)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
