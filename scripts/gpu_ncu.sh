#!/bin/bash
# Round-2 ncu evidence: launch lists (inference step x3, one encoder training step) and --set full captures
TAG=${1:-r02k}
mkdir -p gpurun_out
FBKST_PROFILE_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list exit=$? lines=$(wc -l < gpurun_out/${TAG}_launches.csv)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file gpurun_out/${TAG}_train_launches.csv python scripts/profile_train_step.py > gpurun_out/${TAG}_train_launches.log 2>&1
echo "train launch list exit=$? lines=$(wc -l < gpurun_out/${TAG}_train_launches.csv)"
cap() { # name regex skip count script
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$1 python $5 > gpurun_out/$1.log 2>&1
  echo "$1 exit=$? $(ls -la gpurun_out/$1.ncu-rep 2>/dev/null | awk '{print $5}')"
}
cap ${TAG}_attention 'attention_fwd' 0 1 scripts/profile_step.py
cap ${TAG}_gemm2 'gemm2_kernel' 1 4 scripts/profile_step.py
cap ${TAG}_attn_train 'attn_train_kernel' 0 3 scripts/profile_train_step.py
