#!/bin/bash
# A/B of an environment switch on the bench: usage gpu_ab.sh VAR
mkdir -p gpurun_out
bash scripts/gpu_tests.sh 2>&1 | grep -v "^===" | tee gpurun_out/tests_summary.txt
for v in 1 0; do
  env $1=$v timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python - <<PY
import json
try:
    r = json.load(open("gpurun_out/bench_ab_$v.json"))
    print("$1=$v value %.3fM e2e %.3fM ms/step %.4f" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"]))
except Exception as e:
    print("$1=$v failed", e); print(open("gpurun_out/bench_ab_$v.err").read()[-2000:])
PY
done
