"""Encoder training step (train(): BatchNorm batch statistics, every dropout site, hand-written backward) at the
cfg2 shape, bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off` launch lists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from fbkst_b200.config import build_encoder  # noqa: E402

cfg = bench.CONFIGS["cfg2"]
model, lengths = cfg["model"], cfg["lengths"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc = build_encoder(model, None, device="cpu")
bench.randomise_norm_stats(enc, 1)
enc = enc.to(dev).train()
for p in enc.parameters():
    p.requires_grad_(True)
B, T = len(lengths), max(lengths)
L = ((T + 1) // 2 + 1) // 2
plan = bench.label_plan(L, B, model["vocab"], seed=7).to(dev)
enc.ctc_logit_bump = (plan.to(torch.int32).contiguous(), bench.CTC_MARGIN)
x, l = bench.make_batch(lengths, model["feat_dim"], 1234)
x, l = x.to(dev), l.to(dev)


def step():
    out = enc(x, l, return_all_hiddens=True)
    loss = out.encoder_out.float().pow(2).mean() + 1e-3 * out.ctc_out.float().pow(2).mean()
    loss.backward()
    enc.zero_grad(set_to_none=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
