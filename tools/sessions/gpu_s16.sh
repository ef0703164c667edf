#!/bin/bash
# session 16: SDC path, finalize step of a finished lane deferred to the next round (A/B within one call), then the GPU tests of the SDC path
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in nodefer defer; do
  f=build/variants/libnyx_hc_$v.so
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 struct 3 2>&1 | grep " rep " | tail -3
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 128 5 struct 3 2>&1 | grep " rep " | tail -2
done
done
echo "== vec (default lib)"
timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep " | tail -2
} > gpurun_out/s16.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s16.log
cat gpurun_out/s16.log; tail -5 gpurun_out/s16_pytest.log
