#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh 2>&1 | tee gpurun_out/tests_summary.txt
timeout 300 python scripts/bench_small.py 10 2>&1 | tee gpurun_out/bench_small.txt
timeout 300 python scripts/bench_attn.py 10 2>&1 | tee gpurun_out/bench_attn.txt
