#!/bin/bash
# A/B of the training step (cfg4) under environment variants: scripts/ab_cfg4.sh "VAR=a VAR=b ..." [reps] [steps]
variants=${1:-"FBKST_PRETRANSPOSE=fwd FBKST_PRETRANSPOSE=bwd FBKST_PRETRANSPOSE=off"}
reps=${2:-3}
steps=${3:-10}
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
  for v in $variants; do
    env $v python bench.py --config cfg4 --steps $steps --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['train_step_split'];print('$v', 'rep$r', d['ms_per_step'], d['e2e']['ms_per_step'], s['forward_and_loss_ms'], s['backward_and_allreduce_ms'])"
  done
done | tee gpurun_out/ab_cfg4.txt
