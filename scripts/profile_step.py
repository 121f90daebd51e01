"""One bench-shaped step (BASELINE configs[1]) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off -k regex:<kernel> -c N` captures.  usage: profile_step.py [cfg2|cfg1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from fbkst_b200 import ops  # noqa: E402
from fbkst_b200.config import build_encoder  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = bench.CONFIGS[name]
model, lengths = cfg["model"], cfg["lengths"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc = build_encoder(model, None, device="cpu")
bench.randomise_norm_stats(enc, 1)
enc = enc.to(dev).eval()
B, T = len(lengths), max(lengths)
L = ((T + 1) // 2 + 1) // 2
plan = bench.label_plan(L, B, model["vocab"], seed=7).to(dev)
enc.ctc_logit_bump = (plan.to(torch.int32).contiguous(), bench.CTC_MARGIN)  # as bench.py (fused ctc_fc epilogue)
x, l = bench.make_batch(lengths, model["feat_dim"], 1234)
x = x.to(dev)
len32 = torch.tensor(lengths, dtype=torch.int32, device=dev)
for _ in range(3):
    enc(ops.cmvn(x, len32), l)
torch.cuda.synchronize()
steps = int(os.environ.get("FBKST_PROFILE_STEPS", "1"))
torch.cuda.profiler.start()
for _ in range(steps):
    out = enc(ops.cmvn(x, len32), l)
ops.cmvn(x, len32)  # closes the last step for scripts/launch_summary.py (steps are delimited by cmvn_stats)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", out.src_lengths.sum().item())
