#!/bin/bash
# attention A/B: parity (attention + encoder), micro-benchmark with the decoupled / split-KV kernel, timeline
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 600 python -m pytest "$@" -x -q > gpurun_out/$name.log 2>&1; echo "$name exit=$? :: $(tail -n 1 gpurun_out/$name.log)"; }
FBKST_ATTN_DEC=1 run attn tests/test_gpu_ops.py -k "attention"
FBKST_ATTN_DEC=1 run encoder tests/test_gpu_encoder.py
grep -E "^(E |FAILED|ERROR)|assert|Error|watchdog" gpurun_out/attn.log gpurun_out/encoder.log | head -n 20
echo "== decoupled"; FBKST_ATTN_DEC=1 timeout 300 python scripts/bench_attn.py 10 2>&1 | tee gpurun_out/bench_attn_dec.txt
echo "== split-KV";  timeout 300 python scripts/bench_attn.py 10 2>&1 | tee gpurun_out/bench_attn_split.txt
FBKST_NVCC_FLAGS="-DFBKST_ATTN_TRACE" python fbk-fairseq-st_b200/build.py > /dev/null && FBKST_ATTN_DEC=1 timeout 300 python scripts/trace_attn.py > gpurun_out/attn_trace_dec.txt 2>&1
python fbk-fairseq-st_b200/build.py > /dev/null
head -n 26 gpurun_out/attn_trace_dec.txt | cut -c1-100
