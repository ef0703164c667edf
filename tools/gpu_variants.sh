#!/bin/bash
# times every library variant under build/variants/ on a 256^3 vec step and a 128^3 struct step
mkdir -p gpurun_out
for f in nyx_b200/csrc/libnyx_hc.so build/variants/*.so; do
  echo "== $f"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 3 vec 3 2>&1 | tail -2
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 128 3 struct 3 2>&1 | tail -1
done 2>&1 | tee gpurun_out/variants.log
