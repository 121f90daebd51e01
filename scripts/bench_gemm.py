"""Per-shape timing of the tcgen05 linear kernel (CUDA events, L2 flushed between launches)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
from fbkst_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
SHAPES = [("qkv", 24000, 1536, 512, False, False, torch.bfloat16),
          ("out_proj", 24000, 512, 512, False, True, torch.float32),
          ("fc1", 24000, 2048, 512, True, False, torch.bfloat16),
          ("fc2", 24000, 512, 2048, False, True, torch.float32),
          ("ctc_fc", 24000, 8005, 512, False, False, torch.bfloat16),
          ("fc3", 24000, 512, 640, True, False, torch.float32),
          ("qkv_post", 6100, 1536, 512, False, False, torch.bfloat16),
          ("fc2_post", 6100, 512, 2048, False, True, torch.float32),
          ("square8k", 8192, 8192, 8192, False, False, torch.bfloat16)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
only = sys.argv[2] if len(sys.argv) > 2 else None
for name, M, N, K, relu, resid, odt in SHAPES:
    if only and only != name:
        continue
    a = torch.randn(M, K, device=d).bfloat16()
    w = (torch.randn(N, K, device=d) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=d)
    res = torch.randn(M, N, device=d) if resid else None
    ts = []
    for i in range(reps + 2):
        flush.fill_(i)
        flush.view(torch.int32).sum()  # read pass: leaves CLEAN lines in L2 (no write-back under the kernel)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = ops.linear(a, w, bias, relu=relu, residual=res, out_dtype=odt)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts = sorted(ts[2:])
    med = ts[len(ts) // 2]
    by = (M * K + N * K) * 2 + M * N * (4 if odt == torch.float32 else 2) + (M * N * 4 if resid else 0)
    print("%-9s M=%5d N=%4d K=%4d  %8.1f us  %7.1f TF/s  %6.0f GB/s  (min %.1f us)" %
          (name, M, N, K, med * 1e3, 2.0 * M * N * K / med / 1e9, by / med / 1e6, ts[0] * 1e3), flush=True)
    del a, w, res, out
