#!/bin/bash
# Build a variant of the library HERE (no GPU needed) for A/B runs on the GPU box:
#   scripts/build_variant.sh NAME FILE.cu "-DFOO=1 ..."   -> fbk-fairseq-st_b200/build/variants/NAME.so
# (all other objects come from the regular build; build/ is git-ignored but travels with gpurun)
set -e
cd "$(dirname "$0")/../fbk-fairseq-st_b200"
python build.py > /dev/null
mkdir -p build/variants
name=$1; src=$2; flags=$3
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
  -c csrc/$src -o build/variants/$name.o
objs=$(ls build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o build/variants/$name.so $objs build/variants/$name.o -gencode arch=compute_100a,code=sm_100a
echo build/variants/$name.so
