#!/bin/bash
# r01g GPU visit: the driver's own test command, smoke, bench (ours graph / eager + reference arm), ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.txt 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/gpu_tests.txt)"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.txt)"
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "bench eager exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit=$?"
python - <<'PY'
import json
for f in ("bench_full", "bench_eager"):
    try:
        r = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.3fM e2e %.3fM ms/step %.4f e2e ms %.4f launches %d single %s" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["e2e"]["ms_per_step"], r["gpu_launches"], r["single_forward_ms"]["median"]))
        print(" roofline", r["roofline"]["kernel"], r["roofline"]["achieved"], r["roofline"]["frac"], "clocks", r["clocks"])
        if f == "bench_full":
            print(" cpu_baseline", r.get("cpu_baseline"))
            for k, v in r["kernels"].items():
                print("  %-50s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k[:50], v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
    except Exception as e:
        print("no bench json", f, e); print(open("gpurun_out/%s.err" % f).read()[-3000:])
PY
