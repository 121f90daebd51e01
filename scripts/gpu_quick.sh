#!/bin/bash
# parity suite + bench (no CPU baseline) -- the inner loop of kernel work
mkdir -p gpurun_out
bash scripts/gpu_tests.sh 2>&1 | tee gpurun_out/tests_summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/bench_quick.json"))
    print("value %.3fM e2e %.3fM ms/step %.4f launches %d" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["gpu_launches"]))
    for k, v in r["kernels"].items():
        print("  %-16s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
except Exception as e:
    print("no bench json", e); print(open("gpurun_out/bench_quick.err").read()[-3000:])
PY
