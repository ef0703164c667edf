#!/bin/bash
mkdir -p gpurun_out
{
for v in r1kernel default timing2; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep -v "^$"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 128 5 struct 3 2>&1 | grep -v "^$"
  if [ $v != timing2 ]; then HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 3 2>&1 | grep " rep "; fi
done
echo "== carve-out sweep (default build)"
for kb in 132 164 196 228; do
  echo "-- NYX_HC_CARVEOUT_KB=$kb"
  NYX_HC_CARVEOUT_KB=$kb timeout 300 python tools/prof_driver.py 256 4 vec 3 2>&1 | grep " rep " | tail -2
  NYX_HC_CARVEOUT_KB=$kb timeout 300 python tools/prof_driver.py 128 4 struct 3 2>&1 | grep " rep " | tail -2
done
} > gpurun_out/s5.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5.log
tail -3 gpurun_out/s5_pytest.log
