#!/bin/bash
# attention work loop: parity (attention + encoder), micro-benchmark, timeline trace, quick bench
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -x -q > gpurun_out/$name.log 2>&1; echo "$name exit=$? :: $(tail -n 1 gpurun_out/$name.log)"; }
run attn tests/test_gpu_ops.py -k "attention"
run encoder tests/test_gpu_encoder.py
grep -E "^(E |FAILED|ERROR)|assert|Error" gpurun_out/attn.log gpurun_out/encoder.log | head -n 20
timeout 300 python scripts/bench_attn.py 10 2>&1 | tee gpurun_out/bench_attn.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/bench_quick.json"))
    print("value %.3fM e2e %.3fM ms/step %.4f launches %d" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["gpu_launches"]))
    for k, v in r["kernels"].items():
        print("  %-16s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k[:40], v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
except Exception as e:
    print("no bench json", e); print(open("gpurun_out/bench_quick.err").read()[-3000:])
PY
# timeline (rebuild with the trace hooks, then restore)
FBKST_NVCC_FLAGS="-DFBKST_ATTN_TRACE" python fbk-fairseq-st_b200/build.py > /dev/null && timeout 300 python scripts/trace_attn.py 2>&1 | tee gpurun_out/attn_trace.txt
python fbk-fairseq-st_b200/build.py > /dev/null
