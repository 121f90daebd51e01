"""Eager loop over the rotating bench batches of a ragged configuration: first step / first state with non-finite values."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
import bench  # noqa: E402
from fbkst_b200 import ops  # noqa: E402
from fbkst_b200.config import build_encoder  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
cfg = bench.CONFIGS[name]
model, lengths = cfg["model"], cfg["lengths"]
dev = torch.device("cuda", 0)
B, T, Fd = len(lengths), max(lengths), model["feat_dim"]
torch.manual_seed(0)
enc = build_encoder(model, None, device="cpu")
bench.randomise_norm_stats(enc, 1)
enc = enc.to(dev).eval()
enc.use_cuda_graph = False
L = ((T + 1) // 2 + 1) // 2
plan = bench.label_plan(L, B, model["vocab"], seed=7).to(dev)
enc.ctc_logit_bump = (plan.to(torch.int32).contiguous(), bench.CTC_MARGIN)
batches = [(bench.make_batch(lengths, Fd, 1234 + i)[0].to(dev), lengths) for i in range(9)]
len32 = torch.tensor(lengths, dtype=torch.int32, device=dev)
lt = torch.tensor(lengths)
att_len = [((n + 1) // 2 + 1) // 2 for n in lengths]
for k in range(n_steps):
    x, _ = batches[k % 9]
    xn = ops.cmvn(x, len32)
    ws_mode = os.environ.get("WS_MODE") is not None
    o = enc(xn, lt, return_all_hiddens=not ws_mode)
    if os.environ.get("NO_SYNC") is None or k % 50 == 49:
        torch.cuda.synchronize()
    else:
        continue
    if ws_mode and os.environ.get("WS_TRACK") and enc._ws:
        w = enc._ws[0][1]
        mn = max(o.src_lengths.tolist())
        def reg(t, lo, hi):
            v = t.float().view(L, B, -1)[lo:hi]
            return float(v.nan_to_num(0, 0, 0).abs().max()) if v.numel() else 0.0
        t_tile = (mn * B + 255) // 256 * 256 // B
        if k < 12 or k % 8 == 0:
            print("step %3d max_new %d: " % (k, mn) + "  ".join("%s[<%d]=%.3g [%d..%d)=%.3g [>=%d]=%.3g" % (
                nm, mn, reg(w[nm], 0, mn), mn, t_tile + 1, reg(w[nm], mn, t_tile + 1), t_tile + 1, reg(w[nm], t_tile + 1, L))
                for nm in ("x0", "x1", "qkv", "att")), flush=True)
    if ws_mode:
        eo_ = o.encoder_out
        f = ~torch.isfinite(eo_).all(dim=-1)
        if bool(f.any()):
            print("step %d: encoder_out %s non-finite rows per utt: %s" % (k, tuple(eo_.shape), [(b, int(f[:, b].sum())) for b in range(B) if bool(f[:, b].any())]))
            for lane, (key, w) in (enc._ws or {}).items():
                for nm, t in w.items():
                    tf = t.float()
                    nf = ~torch.isfinite(tf)
                    rows = nf.any(dim=-1).nonzero().flatten() if tf.dim() == 2 else nf.reshape(tf.shape[0], -1).any(dim=-1).nonzero().flatten()
                    print("   workspace lane %s %-4s shape %s non-finite %d rows(first) %s (t,b of first: %s) max|finite| %.3g" % (
                        lane, nm, tuple(t.shape), int(nf.sum()), rows[:6].tolist(),
                        [(int(r) // B, int(r) % B) for r in rows[:6].tolist()], float(tf.nan_to_num(0, 0, 0).abs().max())))
            break
        continue
    nl = o.src_lengths.tolist()
    bad_any = False
    if not bool(torch.isfinite(xn).all()):
        print("step %d: cmvn output non-finite" % k); bad_any = True
    for i, st in enumerate(o.encoder_states):
        Ls = st.shape[0]
        pre = i <= model["ctc_layer"] - 1 if model["ctc_layer"] else True
        rows_bad = {}
        for b in range(B):
            n = min(att_len[b] if Ls == L and pre else nl[b], Ls)
            f = ~torch.isfinite(st[:n, b]).all(dim=-1)
            if bool(f.any()):
                rows_bad[b] = (f.nonzero().flatten().tolist()[:4], int(f.sum()), n)
        if rows_bad:
            print("step %d (batch %d) state %d shape %s: %s" % (k, k % 9, i, tuple(st.shape), rows_bad)); bad_any = True
            break
    if bad_any:
        ws = enc._ws
        for lane, (key, w) in (ws or {}).items():
            for nm, t in w.items():
                nf = int((~torch.isfinite(t.float())).sum())
                mx = float(t.float().nan_to_num(0, 0, 0).abs().max())
                print("   workspace lane %s %-4s shape %s non-finite %d max|finite| %.3g" % (lane, nm, tuple(t.shape), nf, mx))
        break
else:
    print("no non-finite value in %d steps" % n_steps)
