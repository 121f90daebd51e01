"""Diagnostic dump for the tcgen05 GEMM (run on the GPU box when test_linear_plain fails)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
from fbkst_b200 import ops  # noqa: E402

d = torch.device("cuda:0")


def run(M, N, K, mode):
    g = torch.Generator().manual_seed(1)
    if mode == "ones":
        a, w = torch.ones(M, K), torch.ones(N, K)
    elif mode == "arow":  # a[m, k] = m  -> out[m, n] = m * K
        a, w = torch.arange(M).float()[:, None].expand(M, K).contiguous() / 64, torch.ones(N, K)
    elif mode == "wrow":
        a, w = torch.ones(M, K), torch.arange(N).float()[:, None].expand(N, K).contiguous() / 64
    elif mode == "kramp":  # distinguishes k positions: a[m,k] = (k==m%K), w[n,k] = k  -> out[m,n] = m%K
        a = torch.zeros(M, K)
        a[torch.arange(M), torch.arange(M) % K] = 1
        w = torch.arange(K).float()[None, :].expand(N, K).contiguous()
    else:
        a, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / math.sqrt(K)
    a, w = a.bfloat16().to(d), w.bfloat16().to(d)
    ref = a.float() @ w.float().t()
    out = ops.linear(a, w, None, out_dtype=torch.float32)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    print("M=%d N=%d K=%d %-6s max_err=%.4g ref_max=%.4g bad=%d/%d" %
          (M, N, K, mode, err.max().item(), ref.abs().max().item(),
           int((err > 1e-2 * ref.abs().max()).sum()), M * N))
    if err.max() > 1e-2 * ref.abs().max():
        bad = (err > 1e-2 * ref.abs().max())
        rows = bad.any(1).nonzero().flatten()[:16].tolist()
        cols = bad.any(0).nonzero().flatten()[:16].tolist()
        print("   first bad rows", rows, "cols", cols)
        print("   out[0:4,0:8]", out[0:4, 0:8].tolist())
        print("   ref[0:4,0:8]", ref[0:4, 0:8].tolist())


for shape in [(128, 128, 64), (128, 128, 128), (128, 256, 64), (256, 128, 256), (128, 512, 512)]:
    for mode in ["ones", "arow", "wrow", "kramp", "rand"]:
        try:
            run(*shape, mode)
        except Exception as e:  # noqa: BLE001
            print("EXC", shape, mode, repr(e))
            sys.exit(1)
