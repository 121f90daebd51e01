#!/bin/bash
# pipeline work loop: encoder parity (incl. pipeline lanes), quick bench, e2e probe, attention timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py -x -q > gpurun_out/encoder.log 2>&1; echo "encoder exit=$? :: $(tail -n 1 gpurun_out/encoder.log)"
grep -E "^(E |FAILED|ERROR)|assert|Error" gpurun_out/encoder.log | head -n 20
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/bench_quick.json"))
    print("value %.3fM e2e %.3fM ms/step %.4f e2e ms %.4f launches %d" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["e2e"]["ms_per_step"], r["gpu_launches"]))
except Exception as e:
    print("no bench json", e); print(open("gpurun_out/bench_quick.err").read()[-3000:])
PY
FBKST_NVCC_FLAGS="-DFBKST_ATTN_TRACE" python fbk-fairseq-st_b200/build.py > /dev/null && timeout 300 python scripts/trace_attn.py > gpurun_out/attn_trace.txt 2>&1
python fbk-fairseq-st_b200/build.py > /dev/null
head -n 30 gpurun_out/attn_trace.txt
