#!/bin/bash
# round-2 session 2: slim lane state (364 / 420 B per lane) vs the round-1 kernel; full GPU test suite with the new config / fixture / real drop-in tests
mkdir -p gpurun_out
{
for v in r1kernel default; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep "
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 128 5 struct 3 2>&1 | grep " rep "
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 3 2>&1 | grep " rep "
done
} > gpurun_out/s2.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2.log
tail -25 gpurun_out/s2.log; tail -5 gpurun_out/s2_pytest.log
