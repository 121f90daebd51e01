"""Per-kernel device time of the cfg4 training step (torch.profiler; informative, not a bench number)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py", "--config", "cfg4", "--steps", "2", "--warmup", "3"]
import torch
from torch.profiler import profile, ProfilerActivity
import bench

_orig = bench.run_train


def patched(args, rank, world, local_rank):
    # run the bench's own construction, but profile its timed steps
    import types
    src = _orig
    return src(args, rank, world, local_rank)


with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    args = type("A", (), dict(config="cfg4", steps=2, warmup=3, clock_interval_ms=0))()
    r = bench.run_train(args, 0, 1, 0)
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type.name == "CUDA":
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total device time %.1f ms over %d profiled steps (warm-up, timed, e2e and attribution passes)" % (tot / 1e3, 3 + 2 + 2 + 3))
for t, n, k in rows[:45]:
    print("%8.2f ms %5.1f%%  n=%-5d %s" % (t / 1e3, 100 * t / tot, n, k[:110]))
