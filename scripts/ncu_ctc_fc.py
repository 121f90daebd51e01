import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
import torch
from fbkst_b200 import ops
L, B, V, K = 375, 64, 8005, 512
M = L * B
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
w = (torch.randn(V, K, generator=g) / 22).to(torch.bfloat16).cuda()
bias = torch.randn(V, generator=g).cuda()
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
for _ in range(2):
    ops.linear(a, w, bias)
    ops.linear(a, w, bias, out_dtype=torch.float32)
    ops.linear_argmax(a, w, bias, lens, L, B, want_prob=False)
torch.cuda.synchronize()
