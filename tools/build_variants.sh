#!/bin/bash
# builds experiment variants of the CUDA library into build/variants/ (travels to the GPU box, stays out of git)
# usage: build_variants.sh name:"-DHC_X=1 -DHC_Y=2" ...
set -e
cd "$(dirname "$0")/.."
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared"
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc $FLAGS $defs nyx_b200/csrc/nyx_hc.cu -o build/variants/libnyx_hc_$name.so -Xptxas -v > build/variants/$name.log 2>&1 || { grep -m5 "error" build/variants/$name.log; echo "BUILD FAILED: $name"; exit 1; }
  grep -A1 "Compiling.*\(sorted\|hc_flow\|hc_integrate\)" build/variants/$name.log | grep -E "Used|spill" | tr '\n' ' ' || true
  echo " <- $name"
done
