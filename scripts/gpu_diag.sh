#!/bin/bash
# conv1 CTAs/SM sweep + repeated bench with e2e diagnostics + cgroup CPU throttling counters
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
for ps in 3 4 5 6; do echo "per_sm=$ps"; FBKST_CONV1_PER_SM=$ps timeout 120 python scripts/bench_small.py 10 conv1 2>&1 | grep conv1; done
timeout 120 python scripts/bench_small.py 10 cmvn 2>&1 | grep cmvn
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | grep -i thrott
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(r['value']/1e6,2), round(r['e2e']['value']/1e6,2), r['ms_per_step'], r['e2e']['ms_per_step'], r['e2e']['result_to_result_ms']['max'], r['e2e']['result_to_result_ms']['first_result_ms'], r['single_forward_ms']['max'], r['kernels']['conv1_relu_bn']['ms_per_step'], r['kernels']['cmvn']['ms_per_step'])"; done
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | grep -i thrott
