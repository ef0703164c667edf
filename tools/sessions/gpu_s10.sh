#!/bin/bash
mkdir -p gpurun_out
{
for v in fine2 default; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep " | tail -3
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 3 2>&1 | grep " rep " | tail -2
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 6 2>&1 | grep " rep " | tail -2
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 vec 2 2>&1 | grep " rep " | tail -2
done
} > gpurun_out/s10.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10.log
cat gpurun_out/s10.log; tail -3 gpurun_out/s10_pytest.log
