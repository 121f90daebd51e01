#!/bin/bash
# conv1_tc: shared-memory carve-out sweep, alone (micro-benchmark) and inside a bench step (kernel table)
for cv in -1 100; do
  echo "== carveout=$cv"
  FBKST_CONV1_CARVEOUT=$cv timeout 120 python scripts/bench_small.py 10 conv1 2>&1 | grep "conv1"
  FBKST_CONV1_CARVEOUT=$cv timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  bench: ms/step', r['ms_per_step'], 'e2e', r['e2e']['ms_per_step'], 'conv1', r['kernels']['conv1_relu_bn']['ms_per_step'])"
done
