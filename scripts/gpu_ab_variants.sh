#!/bin/bash
# A/B over prebuilt library variants (scripts/build_variant.sh): usage gpu_ab_variants.sh "<cmd>" name...
mkdir -p gpurun_out
LIB=fbk-fairseq-st_b200/fbkst_b200/libfbkst_b200.so
cp $LIB /tmp/lib_orig.so
cmd=$1; shift
for v in "$@"; do
  cp fbk-fairseq-st_b200/build/variants/$v.so $LIB
  echo "=== variant $v"
  bash -c "$cmd" 2>&1 | tee gpurun_out/ab_$v.txt
done
cp /tmp/lib_orig.so $LIB
