#!/bin/bash
# A/B of the inference bench under environment variants: scripts/gpu_ab_env.sh "VAR=a VAR=b" [reps] [steps]
variants=${1:-"FBKST_CTC_FUSED=1 FBKST_CTC_FUSED=0"}
reps=${2:-2}
steps=${3:-20}
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
  for v in $variants; do
    env $v python bench.py --steps $steps --warmup 5 --no-cpu-baseline --soak-seconds 0 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v', 'rep$r', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'single', d['single_forward_ms']['median'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done | tee gpurun_out/ab_env.txt
