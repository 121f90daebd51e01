#!/bin/bash
# Round-2 GPU visit: whole -m gpu suite (per file, own process + timeout), then the bench line.
# usage: gpurun -- bash scripts/gpu_r2.sh [tag] [bench: 0|1]
TAG=${1:-r02}
BENCH=${2:-1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
rm -f gpurun_out/parity_full.json
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -x -q -m gpu > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n exit=$? :: $(tail -n 1 gpurun_out/${TAG}_$n.log)"
done
for f in tests/test_gpu_*.py; do n=$(basename $f .py); echo "=== $n"; grep -E "^(E |FAILED|ERROR)|Error" gpurun_out/${TAG}_$n.log | head -n 14; done
[ -f gpurun_out/parity_full.json ] && cat gpurun_out/parity_full.json
if [ "$BENCH" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
  tail -c 1500 gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    r = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value %.3fM e2e %.3fM ms/step %.4f launches %d" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["gpu_launches"]))
    rf = r["roofline"]
    print("roofline", rf["kernel"], rf["achieved"], rf["frac"], "whole", rf["whole_step"], "sustained", rf["sustained"])
    print("parity", r.get("parity")); print("cpu", r.get("cpu_baseline")); print("ref gpu eager", r.get("reference_gpu_eager"))
    for k, v in r["kernels"].items():
        print("  %-80s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
except Exception as e:
    print("no bench json", e)
PY
fi
