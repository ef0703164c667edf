#!/bin/bash
# builds experiment variants of the CUDA library into build/variants/ (travels to the GPU box, stays out of git)
set -e
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared"
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc $FLAGS $defs nyx_b200/csrc/nyx_hc.cu -o build/variants/libnyx_hc_$name.so -Xptxas -v 2>&1 | grep -E "hc_integrate_kernelILi0|Used|spill" | grep -A2 "ILi0EEE" | grep -E "Used|spill" | head -2 | tr '\n' ' '
  echo " <- $name"
done
