#!/bin/bash
# front-end work loop (conv1 v2, fc3 coalesced epilogue): parity, micro-benchmark A/B, quick bench with kernel table
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -x -q > gpurun_out/$name.log 2>&1; echo "$name exit=$? :: $(tail -n 1 gpurun_out/$name.log)"; }
run ops tests/test_gpu_ops.py -k "conv or linear or fc3"
run encoder tests/test_gpu_encoder.py
grep -E "^(E |FAILED|ERROR)|assert|Error|watchdog" gpurun_out/ops.log gpurun_out/encoder.log | head -n 20
echo "== conv1 tcgen05"; timeout 120 python scripts/bench_small.py 10 conv 2>&1 | tail -n 3
echo "== conv1 SIMT"; FBKST_CONV1_SIMT=1 timeout 120 python scripts/bench_small.py 10 conv1 2>&1 | tail -n 2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/bench_quick.json"))
    print("value %.3fM e2e %.3fM ms/step %.4f e2e ms %.4f" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["e2e"]["ms_per_step"]))
    for k, v in r["kernels"].items():
        print("  %-50s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k[:50], v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
except Exception as e:
    print("no bench json", e); print(open("gpurun_out/bench_quick.err").read()[-3000:])
PY
