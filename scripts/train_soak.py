"""Training soak: N optimisation steps of the encoder alone (train(): batch statistics, every dropout site,
hand-written backward, Adam) on rotating ragged batches; checks for hangs, non-finite values and that the loss
goes down.  usage: train_soak.py [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from fbkst_b200 import criterion as C  # noqa: E402
from fbkst_b200.config import build_encoder  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
model = dict(bench.CONFIGS["cfg2"]["model"], layers=6, ctc_layer=4, vocab=1005)
torch.manual_seed(0)
enc = build_encoder(model, None, device="cpu").cuda().train()
for p in enc.parameters():
    p.requires_grad_(True)
opt = torch.optim.Adam(enc.parameters(), lr=2e-4, fused=True)
g = torch.Generator().manual_seed(1)
batches = []
for i in range(6):
    lens = sorted(torch.randint(300, 1500, (24,), generator=g).tolist(), reverse=True)
    x, l = bench.make_batch(lens, 40, 100 + i)
    U = 20
    tgt = torch.randint(4, 1003, (len(lens), U), generator=g)
    batches.append((x.cuda(), l.cuda(), tgt.cuda(), torch.full((len(lens),), U, dtype=torch.long).cuda()))
losses, t0 = [], time.time()
for s in range(steps):
    x, l, tgt, tl = batches[s % len(batches)]
    torch.manual_seed(1000 + s)
    out = enc(x, l, return_all_hiddens=True)
    mask = out.ctc_padding_mask
    ctc, totals, _ = C.ctc_loss_train(out.ctc_out, None if mask is None else mask.t(), tgt, tl, 1004)
    reg = out.encoder_out.float().pow(2).mean()
    loss = ctc / x.shape[0] + reg
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    if s % 25 == 0 or s == steps - 1:
        v = loss.item()
        assert v == v and abs(v) < 1e9, "non-finite loss at step %d" % s
        losses.append(v)
        print("step %4d  loss %.4f  (ctc/utt %.3f)  %.1f s" % (s, v, ctc.item() / x.shape[0], time.time() - t0), flush=True)
torch.cuda.synchronize()
assert all(torch.isfinite(p).all() for p in enc.parameters()), "non-finite parameter"
assert losses[-1] < 0.8 * losses[0], "loss did not go down: %s" % losses
print("soak ok: %d steps, loss %.3f -> %.3f, %.1f s, %.1f ms/step" %
      (steps, losses[0], losses[-1], time.time() - t0, (time.time() - t0) / steps * 1e3))
