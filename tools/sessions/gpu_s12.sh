#!/bin/bash
mkdir -p gpurun_out
{
for v in preskip sync2 default; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 6 vec 3 2>&1 | grep " rep " | tail -4
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 struct 3 2>&1 | grep " rep " | tail -3
done
} > gpurun_out/s12.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12.log
cat gpurun_out/s12.log; grep "^\[parity\]" gpurun_out/s12_pytest.log | cut -c1-200
