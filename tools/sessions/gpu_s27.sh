#!/bin/bash
# session 27: per-cell constants saved once per cell instead of every round (A/B within one call), then the SDC / Strang parity tests
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in base savecell; do
  f=build/variants/libnyx_hc_$v.so
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep " | tail -2
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 struct 3 2>&1 | grep " rep " | tail -2
done
done
} > gpurun_out/s27.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s27_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s27.log
cat gpurun_out/s27.log; tail -3 gpurun_out/s27_pytest.log
