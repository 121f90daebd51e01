#!/bin/bash
# ncu --set full of the front-end kernels of one bench-shaped step (r01g: conv1 on tcgen05, float4 CMVN, embed_remap_stats)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k "regex:conv1_tc|conv2_kernel|embed_remap|cmvn_" -c 5 -f -o gpurun_out/r01g_front python scripts/profile_step.py > gpurun_out/r01g_front.log 2>&1
echo "exit=$? $(ls -la gpurun_out/r01g_front.ncu-rep 2>/dev/null | awk '{print $5}')"
