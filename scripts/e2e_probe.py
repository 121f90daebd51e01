#!/usr/bin/env python
"""Where does the e2e (host-buffer) step lose time against the device-resident step?

Measures on one GPU, cfg2:
  A  resident steps back to back (no flush, no delay)           -> sustained device ms/step + clocks
  B  resident steps with the bench's flush + 1 ms delay          -> what bench.py `value` sees
  C  EncoderPipeline.run for 20 / 100 / 300 steps (wall clock)   -> e2e ms/step and its asymptote
  D  per-step CUDA events on the pipeline's compute stream       -> device time + gaps inside e2e
  E  PCIe: pinned H2D of one input batch, D2H of one output      -> copy times alone
  F  host time to enqueue one pipeline step
Prints one JSON object.  Not a bench value: diagnostic only.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from fbkst_b200 import ops  # noqa: E402
from fbkst_b200.config import build_encoder  # noqa: E402
from fbkst_b200.pipeline import EncoderPipeline  # noqa: E402


def smi():
    try:
        o = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw",
                            "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
        return o.strip()
    except OSError:
        return None


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = bench.CONFIGS["cfg2"]
    model, lengths = cfg["model"], cfg["lengths"]
    B, T, Fd = len(lengths), max(lengths), model["feat_dim"]
    torch.manual_seed(0)
    enc = build_encoder(model, None, device="cpu")
    bench.randomise_norm_stats(enc, 1)
    enc = enc.to(dev).eval()
    enc.use_cuda_graph = True
    L = ((T + 1) // 2 + 1) // 2
    plan = bench.label_plan(L, B, model["vocab"], seed=7).to(dev)

    def bump(mod, inp, out):
        out.scatter_add_(2, plan.unsqueeze(-1),
                         torch.full((L, B, 1), bench.CTC_MARGIN, dtype=out.dtype, device=out.device))
    enc.ctc_fc.register_forward_hook(bump)
    host = [bench.make_batch(lengths, Fd, 1234 + i) for i in range(4)]
    host = [(x.pin_memory(), l) for x, l in host]
    devb = [(x.to(dev), l) for x, l in host]
    len32 = torch.tensor(lengths, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pipe = EncoderPipeline(enc, normalize=True, device=dev)
    res = {}

    def resident(i):
        x, l = devb[i % 4]
        return enc(ops.cmvn(x, len32), l)

    for i in range(5):
        resident(i)
    for _ in pipe.run(host[i % 4] for i in range(5)):
        pass
    torch.cuda.synchronize()

    # A: back to back
    for n in (20, 200):
        torch.cuda._sleep(int(2e-3 * 1.9e9))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n):
            resident(i)
        e.record()
        torch.cuda.synchronize()
        res["A_resident_back_to_back_%d" % n] = dict(ms_per_step=s.elapsed_time(e) / n, smi=smi())

    # B: the bench's timed region 1
    evs = []
    for i in range(20):
        flush.fill_(i & 0xFF)
        torch.cuda._sleep(bench.LOOKAHEAD_CYCLES)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        resident(i)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in evs)
    res["B_flush_delay"] = dict(min=ms[0], median=ms[len(ms) // 2], max=ms[-1])
    # B2: flush but no delay (the pre-lookahead bench)
    evs = []
    for i in range(20):
        flush.fill_(i & 0xFF)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        resident(i)
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in evs)
    res["B2_flush_no_delay"] = dict(min=ms[0], median=ms[len(ms) // 2], max=ms[-1])

    # C: pipeline wall clock
    for n in (20, 100, 300):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in pipe.run(host[i % 4] for i in range(n)):
            pass
        torch.cuda.synchronize()
        res["C_e2e_%d" % n] = dict(ms_per_step=(time.perf_counter() - t0) * 1e3 / n, smi=smi())

    # D: events on the compute stream inside the pipeline (monkey-patched cmvn start / step end)
    marks = []
    orig_cmvn = ops.cmvn

    def cmvn_marked(x, l):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(ev)
        return orig_cmvn(x, l)
    ops.cmvn = cmvn_marked
    import fbkst_b200.pipeline as pl
    pl.ops.cmvn = cmvn_marked
    t_host = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in pipe.run(host[i % 4] for i in range(60)):
        t_host.append(time.perf_counter())
    torch.cuda.synchronize()
    starts = [marks[0].elapsed_time(m) for m in marks]
    d = [b - a for a, b in zip(starts, starts[1:])]
    d.sort()
    res["D_compute_stream_start_to_start_ms"] = dict(min=d[0], median=d[len(d) // 2], max=d[-1])
    hy = sorted(b - a for a, b in zip(t_host, t_host[1:]))
    res["D_host_yield_to_yield_ms"] = dict(min=hy[0] * 1e3, median=hy[len(hy) // 2] * 1e3, max=hy[-1] * 1e3)
    ops.cmvn = orig_cmvn
    pl.ops.cmvn = orig_cmvn

    # E: copies alone
    x_host = host[0][0]
    x_dev = torch.empty_like(devb[0][0])
    out = resident(0).encoder_out
    o_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    for name, fn in (("h2d", lambda: x_dev.copy_(x_host, non_blocking=True)),
                     ("d2h", lambda: o_host.copy_(out, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        nbytes = (x_host if name == "h2d" else out).numel() * 4
        t = s.elapsed_time(e) / 10
        res["E_" + name] = dict(ms=t, MB=nbytes / 1e6, GBs=nbytes / t / 1e6)

    # F: host enqueue cost of one resident step with the GPU kept busy (pure CPU time)
    torch.cuda._sleep(int(20e-3 * 1.9e9))
    t0 = time.perf_counter()
    for i in range(5):
        resident(i)
    res["F_host_enqueue_ms_per_resident_step"] = (time.perf_counter() - t0) * 1e3 / 5
    torch.cuda.synchronize()
    res["cpus"] = os.cpu_count()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
