"""Repro helper: the attention calls of bench cfg1 (L=250, B=8, H=4, no penalty), before / after compression."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from fbkst_b200 import ops  # noqa: E402
from test_gpu_ops import attn_ref  # noqa: E402

d = torch.device("cuda:0")
L, B, H = 250, 8, 4
torch.manual_seed(0)
for name, lens, lim in [("pre", [250, 238, 225, 200, 175, 150, 125, 100], None),
                        ("post", [84, 80, 75, 67, 59, 50, 42, 34], 84),
                        ("post129", [129, 80, 75, 67, 59, 50, 42, 1], 129)]:
    for pen in (False, True):
        worst, nans = 0.0, 0
        for rep in range(20):
            qkv = (torch.randn(L * B, 3 * H * 64, device=d) * 0.7).bfloat16()
            lengths = torch.tensor(lens, dtype=torch.int32, device=d)
            out = torch.full((L * B, H * 64), 7.0, dtype=torch.bfloat16, device=d)
            ql = None if lim is None else torch.tensor([lim], dtype=torch.int32, device=d)
            ops.attention(qkv, lengths, L, B, H, pen, out=out, q_limit=ql)
            torch.cuda.synchronize()
            o = out.float().view(L, B, H * 64)
            ref = attn_ref(qkv, lengths, L, B, H, pen).view(L, B, H * 64)
            for b, n in enumerate(lens):
                if not torch.isfinite(o[:n, b]).all():
                    nans += 1
                    continue
                e = ((o[:n, b] - ref[:n, b]).abs().max() / ref[:n, b].abs().max()).item()
                worst = max(worst, e)
        print("%-8s pen=%d  worst rel err %.4g  non-finite utterances %d" % (name, pen, worst, nans), flush=True)
