#!/bin/bash
# One validation visit: the full GPU suite, smoke(), the default bench line, the UMMA probe
TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.txt 2>&1; echo "pytest exit=$? :: $(tail -n 1 gpurun_out/${TAG}_gpu_tests.txt)"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; echo "smoke exit=$? :: $(tail -n 1 gpurun_out/smoke.txt)"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
timeout 60 fbk-fairseq-st_b200/build/umma_probe > gpurun_out/${TAG}_umma_probe.txt 2>&1
TAG=$TAG python - <<'PY'
import json, os
TAG = os.environ['TAG']
try:
    r = json.load(open("gpurun_out/%s_bench.json" % TAG))
    print("value %.3fM e2e %.3fM ms/step %.4f launches %d" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["gpu_launches"]))
    print(" roofline", r["roofline"]["kernel"], r["roofline"]["achieved"], r["roofline"]["frac"], "parity", r.get("parity"))
    for k, v in r["kernels"].items():
        print("  %-50s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k[:50], v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
except Exception as e:
    print("no bench json", e); print(open("gpurun_out/%s_bench.err" % TAG).read()[-3000:])
PY
