#!/bin/bash
# CTC kernels only: parity (compression + criterion), microbench, quick bench line
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -x -q > gpurun_out/$name.log 2>&1; echo "$name exit=$? :: $(tail -n 1 gpurun_out/$name.log)"; }
run ctc tests/test_gpu_ops.py -k "ctc or custom_ops"
run criterion tests/test_gpu_criterion.py
run encoder tests/test_gpu_encoder.py
for f in ctc criterion encoder; do echo "=== $f"; grep -E "^(E |FAILED|ERROR)|assert|Error" gpurun_out/$f.log | head -n 20; done
timeout 300 python scripts/bench_small.py 10 ctc 2>&1 | tee gpurun_out/bench_small.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_quick.json"))
print("value %.3fM e2e %.3fM ms/step %.4f" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"]))
PY
