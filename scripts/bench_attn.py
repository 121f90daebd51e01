"""Timing of the attention kernel at the cfg2 shapes (CUDA events, L2 flushed between launches)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
from fbkst_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for name, L, B, H, lens in [("pre  L=375", 375, 64, 8, [375] * 64),
                            ("post L=110", 110, 64, 8, [110 - (i % 30) for i in range(64)]),
                            # the in-step shape after CTC compression: worst-case grid, device-side row limit
                            ("post L=375 lim", 375, 64, 8, [120 + (i * 7) % 13 for i in range(64)]),
                            ("long L=1500", 1500, 8, 16, [1500] * 8)]:
    qkv = (torch.randn(L * B, 3 * H * 64, device=d) * 0.7).bfloat16()
    lengths = torch.tensor(lens, dtype=torch.int32, device=d)
    q_limit = torch.tensor([max(lens)], dtype=torch.int32, device=d) if "lim" in name else None
    ts = []
    for i in range(reps + 2):
        flush.fill_(i)
        flush.view(torch.int32).sum()  # read pass: leaves CLEAN lines in L2 (no write-back under the kernel)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = ops.attention(qkv, lengths, L, B, H, True, q_limit=q_limit)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts = sorted(ts[2:])
    med = ts[len(ts) // 2]
    fl = sum(4.0 * n * n * 64 * H for n in lens)
    print("%-15s B=%d H=%d  %8.1f us  %7.1f TF/s (min %.1f us)" % (name, B, H, med * 1e3, fl / med / 1e9, ts[0] * 1e3),
          flush=True)
