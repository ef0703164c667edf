#!/bin/bash
mkdir -p gpurun_out
{
for v in r1kernel default se2 se3 L512 L512se2; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep " | tail -3
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 128 5 struct 3 2>&1 | grep " rep " | tail -3
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 3 2>&1 | grep " rep " | tail -2
done
echo "== parity report se2"
NYX_HC_LIB=$PWD/build/variants/libnyx_hc_se2.so timeout 600 python tests/gpu_report.py 2>&1 | cut -c1-330
} > gpurun_out/s6.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s6_pytest.log 2>&1; echo "pytest default rc=$?" >> gpurun_out/s6.log
NYX_HC_LIB=$PWD/build/variants/libnyx_hc_se2.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/s6_pytest_se2.log 2>&1; echo "pytest se2 rc=$?" >> gpurun_out/s6.log
tail -3 gpurun_out/s6_pytest.log
