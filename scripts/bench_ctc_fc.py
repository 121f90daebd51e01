"""Micro-benchmark of the CTC projection chain at cfg2 (M = 24000 frames, V = 8005, K = 512)."""
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
plan = torch.randint(0, V, (M,), generator=g).to(torch.int32).cuda()


def timeit(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    print("%-58s %8.1f us" % (name, s.elapsed_time(e) / reps * 1e3))


timeit("ctc_fc bf16 logits", lambda: ops.linear(a, w, bias))
timeit("ctc_fc fp32 logits", lambda: ops.linear(a, w, bias, out_dtype=torch.float32))
lg = ops.linear(a, w, bias, out_dtype=torch.float32)
timeit("ctc_argmax on fp32 logits (avg)", lambda: ops.ctc_argmax(lg, lens, L, B, V, False))
timeit("ctc_argmax on fp32 logits (weighted)", lambda: ops.ctc_argmax(lg, lens, L, B, V, True))
timeit("fused ctc_fc + arg-max (avg)", lambda: ops.linear_argmax(a, w, bias, lens, L, B, want_prob=False))
timeit("fused ctc_fc + arg-max (avg, bump)", lambda: ops.linear_argmax(a, w, bias, lens, L, B, want_prob=False, bump=(plan, 30.0)))
timeit("fused ctc_fc + arg-max (weighted, bump)", lambda: ops.linear_argmax(a, w, bias, lens, L, B, want_prob=True, bump=(plan, 30.0)))

# write-only / read-only HBM bandwidth for context (the fp32 logits are a 771 MB pure write)
buf = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
timeit("fill 1 GiB (write only)", lambda: buf.fill_(1))
src = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
timeit("sum 1 GiB (read only)", lambda: src.sum())
dst = torch.empty_like(src)
timeit("copy 1 GiB -> 1 GiB (read + write)", lambda: dst.copy_(src))
