#!/usr/bin/env python
"""Multi-rank e2e diagnostic (torchrun): per rank, without any inter-rank barrier in the timed parts:
  A  EncoderPipeline.run (host buffers) ms/step, 100 steps      B  run_device ms/step
  C  pinned H2D / D2H bandwidth while the other ranks do the same
  D  host yield-to-yield distribution inside A
Prints one JSON line per rank.  Not a bench value."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from fbkst_b200.config import build_encoder  # noqa: E402
from fbkst_b200.pipeline import EncoderPipeline  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    cfg = bench.CONFIGS["cfg2"]
    model, lengths = cfg["model"], cfg["lengths"]
    B, T, Fd = len(lengths), max(lengths), model["feat_dim"]
    torch.manual_seed(0)
    enc = build_encoder(model, None, device="cpu")
    bench.randomise_norm_stats(enc, 1)
    enc = enc.to(dev).eval()
    enc.use_cuda_graph = True
    L = ((T + 1) // 2 + 1) // 2
    plan = bench.label_plan(L, B, model["vocab"], seed=7 + rank).to(dev)

    def bump(mod, inp, out):
        out.scatter_add_(2, plan.unsqueeze(-1),
                         torch.full((L, B, 1), bench.CTC_MARGIN, dtype=out.dtype, device=out.device))
    enc.ctc_fc.register_forward_hook(bump)
    host = [bench.make_batch(lengths, Fd, 1234 + rank * 100 + i) for i in range(9)]
    host = [(x.pin_memory(), l) for x, l in host]
    devb = [(x.to(dev), l) for x, l in host]
    pipe = EncoderPipeline(enc, normalize=True, device=dev)
    for _ in pipe.run(host[i % 9] for i in range(6)):
        pass
    for _ in pipe.run_device(devb[i % 9] for i in range(6)):
        pass
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    res = dict(rank=rank, affinity=len(os.sched_getaffinity(0)), omp=os.environ.get("OMP_NUM_THREADS"),
               torch_threads=torch.get_num_threads())
    n = 100
    for name, fn in (("A_e2e", lambda: pipe.run(host[i % 9] for i in range(n))),
                     ("B_device", lambda: pipe.run_device(devb[i % 9] for i in range(n))),
                     ("A2_e2e", lambda: pipe.run(host[i % 9] for i in range(n)))):
        stamps = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in fn():
            stamps.append(time.perf_counter())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d = sorted(b - a for a, b in zip(stamps, stamps[1:]))
        res[name] = dict(ms_per_step=dt * 1e3 / n, yield_median_ms=d[len(d) // 2] * 1e3, yield_max_ms=d[-1] * 1e3,
                         yield_p90_ms=d[int(0.9 * len(d))] * 1e3)
    x_host = host[0][0]
    x_dev = torch.empty_like(devb[0][0])
    o_dev = torch.empty(100 * 64 * 512, dtype=torch.float32, device=dev)
    o_host = torch.empty(100 * 64 * 512, dtype=torch.float32).pin_memory()
    for name, fn, nb in (("C_h2d", lambda: x_dev.copy_(x_host, non_blocking=True), x_host.numel() * 4),
                         ("C_d2h", lambda: o_host.copy_(o_dev, non_blocking=True), o_dev.numel() * 4)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        res[name] = dict(ms=dt * 1e3, GBs=nb / dt / 1e9)
    # host-side cost of the per-step python (enqueue only; GPU kept busy so nothing blocks on it)
    print(json.dumps(res), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
