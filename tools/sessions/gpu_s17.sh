#!/bin/bash
# session 17: where a round goes with the current kernels (clock64 instrumentation build; shares, not absolutes), both paths
mkdir -p gpurun_out
{
for v in timingN timingL; do
  f=build/variants/libnyx_hc_$v.so
  echo "== $v vec"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 vec 3 2>&1 | tail -45
  echo "== $v struct"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 3 2>&1 | tail -45
done
} > gpurun_out/s17.log 2>&1
tail -5 gpurun_out/s17.log
