"""Where do non-finite values of a ragged bench configuration first appear?  (bench.py set-up, eager forward)"""
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
x, l = bench.make_batch(lengths, Fd, 1234)
len32 = torch.tensor(lengths, dtype=torch.int32, device=dev)
n_soak = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if n_soak:
    from fbkst_b200.pipeline import EncoderPipeline
    enc.use_cuda_graph = os.environ.get("NO_GRAPH") is None
    pipe = EncoderPipeline(enc, normalize=True, device=dev)
    batches = [(bench.make_batch(lengths, Fd, 1234 + i)[0].to(dev), l) for i in range(9)]
    last = None
    for k, o_ in enumerate(pipe.run_device(batches[i % 9] for i in range(n_soak))):
        last = o_
        if k < 40 or k % 500 == 0:
            torch.cuda.synchronize()
            eo_ = o_.encoder_out
            nb = int((~torch.isfinite(eo_)).sum())
            if nb or k % 500 == 0:
                f = ~torch.isfinite(eo_).all(dim=-1)  # [L', B]
                per_utt = [(b, f[:, b].nonzero().flatten().tolist()[:4], int(f[:, b].sum())) for b in range(B) if bool(f[:, b].any())]
                print("pipelined step %d: shape %s lens %s non-finite %d rows per utt (b, first rows, count): %s"
                      % (k, tuple(eo_.shape), o_.src_lengths.tolist(), nb, per_utt), flush=True)
                if nb and k < 40:
                    break
    torch.cuda.synchronize()
    enc.use_cuda_graph = False
xn = ops.cmvn(x.to(dev), len32)
print("cmvn finite:", bool(torch.isfinite(xn).all()))
o = enc(xn, l, return_all_hiddens=True)
torch.cuda.synchronize()
nl = o.src_lengths.tolist()
print("new lengths", nl)
att_len = [((n + 1) // 2 + 1) // 2 for n in lengths]
for i, st in enumerate(o.encoder_states):
    Ls = st.shape[0]
    bad = []
    for b in range(B):
        n = att_len[b] if Ls == L and i < model["ctc_layer"] + 1 else nl[b]
        n = min(n, Ls)
        f = torch.isfinite(st[:n, b]).all(dim=-1)
        if not bool(f.all()):
            rows = (~f).nonzero().flatten().tolist()
            bad.append((b, n, rows[:6], len(rows)))
    print("state %d shape %s non-finite (utt, len, first rows, count): %s" % (i, tuple(st.shape), bad))
eo = o.encoder_out
for b in range(B):
    f = torch.isfinite(eo[:nl[b], b]).all(dim=-1)
    if not bool(f.all()):
        print("encoder_out utt %d len %d bad rows %s" % (b, nl[b], (~f).nonzero().flatten().tolist()[:10]))
