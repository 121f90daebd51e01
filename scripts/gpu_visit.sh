#!/bin/bash
# One full GPU visit: parity suite, bench (graph + eager + full line with cpu baseline), reference arm,
# ncu launch list of the bench command, ncu --set full captures.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu_all.sh
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit=$?"
bash scripts/gpu_ncu_full.sh
