#!/bin/bash
# wide attention kernel visit: parity of the attention ops under FBKST_ATTN_WIDE=1, then the micro-benchmark A/B
mkdir -p gpurun_out
TAG=${1:-wide}
timeout 300 python -m pytest tests/test_gpu_ops.py -k "attention" -x -q > gpurun_out/${TAG}_attn.log 2>&1
echo "attn(wide) exit=$? :: $(tail -n 1 gpurun_out/${TAG}_attn.log)"
grep -E "^(E |FAILED|ERROR)|assert|Error|watchdog" gpurun_out/${TAG}_attn.log | head -n 20
echo "=== round-1 kernels (FBKST_ATTN_WIDE=0)"; FBKST_ATTN_WIDE=0 timeout 200 python scripts/bench_attn.py 10 2>&1 | tee gpurun_out/${TAG}_bench_default.txt
echo "=== wide (default)"; timeout 200 python scripts/bench_attn.py 10 2>&1 | tee gpurun_out/${TAG}_bench_wide.txt
