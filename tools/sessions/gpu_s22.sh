#!/bin/bash
# session 22: (q, qwait) sub-key transposed against the default
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in default subT; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep " | tail -2
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 struct 3 2>&1 | grep " rep " | tail -2
done
done
} > gpurun_out/s22.log 2>&1
cat gpurun_out/s22.log
