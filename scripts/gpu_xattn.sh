#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_xattn.py -x -q > gpurun_out/xattn_tests.txt 2>&1; echo "xattn tests exit=$? :: $(tail -n 1 gpurun_out/xattn_tests.txt)"
grep -E "^(E |FAILED|ERROR)|assert|Error" gpurun_out/xattn_tests.txt | head -n 20
timeout 300 python scripts/bench_small.py 10 xattn 2>&1 | tee gpurun_out/bench_xattn.txt
