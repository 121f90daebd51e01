#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels on small shapes (sanitizers)
mkdir -p gpurun_out
run() { # name, pytest args
  local name=$1; shift
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest "$@" -x -q > gpurun_out/sanitize_$name.log 2>&1
  echo "$name exit=$? :: $(grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitize_$name.log | tail -2 | tr '\n' ' ')"
}
run train_ops tests/test_gpu_train_ops.py -k "dropout or ln_bwd or grad_prep or ctc_compress or bn_train or conv1_wgrad or 64-1-1 or 100-3-2"
run train tests/test_gpu_train.py -k "tiny"
run criterion tests/test_gpu_criterion.py -k "backward and 50-4"
run fused_ctc tests/test_gpu_ops.py -k "argmax_fused and 50-3"
