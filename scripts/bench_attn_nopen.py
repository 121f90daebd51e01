"""bench_attn.py at the cfg2 shape, with and without the log penalty"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
from fbkst_b200 import ops  # noqa: E402
d = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
L, B, H = 375, 64, 8
qkv = (torch.randn(L * B, 3 * H * 64, device=d) * 0.7).bfloat16()
lengths = torch.full((B,), L, dtype=torch.int32, device=d)
for pen in (True, False):
    ts = []
    for i in range(12):
        flush.fill_(i); flush.view(torch.int32).sum()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.attention(qkv, lengths, L, B, H, pen); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts = sorted(ts[2:])
    print("L=375 log_penalty=%s  median %.1f us  min %.1f us" % (pen, ts[len(ts)//2]*1e3, ts[0]*1e3), flush=True)
