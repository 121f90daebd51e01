"""Per-CTA timeline of the attention kernel (debug hook), L=375 B=64 H=8."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fbk-fairseq-st_b200"))
from fbkst_b200 import _lib, ops  # noqa: E402

d = torch.device("cuda:0")
L, B, H = 375, 64, 8
qkv = (torch.randn(L * B, 3 * H * 64, device=d) * 0.7).bfloat16()
lengths = torch.full((B,), L, dtype=torch.int32, device=d)
lib = _lib.load()
lib.fbkst_debug_set_attention_trace.argtypes = [ctypes.c_void_p]
for _ in range(3):
    ops.attention(qkv, lengths, L, B, H, True)
buf = torch.zeros(4096 * 16, dtype=torch.int64, device=d)
assert lib.fbkst_debug_set_attention_trace(buf.data_ptr()) == 0
ops.attention(qkv, lengths, L, B, H, True)
torch.cuda.synchronize()
lib.fbkst_debug_set_attention_trace(None)
t = buf.view(4096, 16)[:1536].cpu()
t0 = t[:, 0].min()
names = ["entry", "setup", "lut", "s_full0", "tile0", "tile1", "tile2", "tile3", "t2_AB", "tile5", "pv_last", "done"]
rel = (t[:, :12] - t[:, :1]).float()
print("per-CTA phase durations (cycles), mean / p90 over 1536 CTAs")
r2 = (t[:, [6 - 1 + 0, 12, 13, 14, 8, 6]] - t[:, 5:6]).float()  # relative to end of tile1
print("tile 2 detail (cycles after end of tile1): s_full %.0f | ld %.0f | max+rescale+pbuf %.0f | A+B %.0f | C+sts+arrive %.0f" %
      (r2[:, 1].mean(), (r2[:, 2] - r2[:, 1]).mean(), (r2[:, 3] - r2[:, 2]).mean(), (r2[:, 4] - r2[:, 3]).mean(),
       (r2[:, 5] - r2[:, 4]).mean()))
prev = None
for i, n in enumerate(names):
    col = rel[:, i]
    dcol = col if prev is None else col - prev
    print("  %-8s at %8.0f   (+%7.0f mean, +%7.0f p90)" % (n, col.mean(), dcol.mean(), dcol.quantile(0.9)))
    prev = col
start = (t[:, 0] - t0).float()
end = (t[:, 11] - t0).float()
print("kernel span %.0f cycles; CTA lifetime mean %.0f; start times quantiles:" % (end.max(), (end - start).mean()),
      [int(start.quantile(q)) for q in (0.1, 0.25, 0.5, 0.75, 0.9)])

