#!/bin/bash
# ncu --set full captures of the kernels of one bench-shaped step (one ncu run per kernel family).
mkdir -p gpurun_out
cap() { # name regex count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k "regex:$2" -c $3 -f -o gpurun_out/$1 python scripts/profile_step.py > gpurun_out/$1.log 2>&1
  echo "$1 exit=$? $(ls -la gpurun_out/$1.ncu-rep 2>/dev/null | awk '{print $5}')"
}
cap r01_gemm2 'gemm2_kernel' 4
cap r01_attention 'attention_fwd' 1
cap r01_ctc 'ctc_(argmax|segment|compress)' 3
cap r01_conv 'conv1_kernel|conv2_kernel|gemm_bf16_kernel' 3
cap r01_small 'layernorm_kernel|cmvn_' 3
