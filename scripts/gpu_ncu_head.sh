#!/bin/bash
# ncu --set full at HEAD: the decoupled attention kernel and the first four CTA-pair GEMM launches of a layer
mkdir -p gpurun_out
cap() { # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$1 python scripts/profile_step.py > gpurun_out/$1.log 2>&1
  echo "$1 exit=$? $(ls -la gpurun_out/$1.ncu-rep 2>/dev/null | awk '{print $5}')"
}
cap r01i_attention 'attention_fwd' 0 1
cap r01i_gemm2 'gemm2_kernel' 1 4
