#!/bin/bash
# times every library variant under build/variants/ (and the default build) on a 256^3 vec step and a 128^3 struct step: best of 6 / 4 repetitions
# (tools/gpu_ab_e2e.sh: A/B of the descriptor staging path on the end-to-end leg of bench.py)
mkdir -p gpurun_out
for f in nyx_b200/csrc/libnyx_hc.so build/variants/*.so; do
  echo "== $f"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 6 vec 3 2>&1 | grep " rep " | sort -t: -k2 -n | head -1
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 128 4 struct 3 2>&1 | grep " rep " | sort -t: -k2 -n | head -1
done 2>&1 | tee gpurun_out/variants.log
