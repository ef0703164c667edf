#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_variants.sh > /dev/null 2>&1; cat gpurun_out/variants.log
for v in a352 a384; do
  NYX_HC_LIB=$PWD/build/variants/libnyx_hc_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "struct or inhomo or full_size or ragged or options or concurrent or host_buffer" > gpurun_out/pytest_$v.log 2>&1; echo "pytest $v rc=$?"; tail -2 gpurun_out/pytest_$v.log
done
for v in a352 a384; do
  echo "== sweep z=6,3 with $v"
  HC_LIB=$PWD/build/variants/libnyx_hc_$v.so timeout 300 python tools/prof_driver.py 256 3 struct 6 2>&1 | grep " rep " | sort -t: -k2 -n | head -1
  HC_LIB=$PWD/build/variants/libnyx_hc_$v.so timeout 300 python tools/prof_driver.py 256 3 struct 2 2>&1 | grep " rep " | sort -t: -k2 -n | head -1
done
echo "== default"
timeout 300 python tools/prof_driver.py 256 3 struct 6 2>&1 | grep " rep " | sort -t: -k2 -n | head -1
timeout 300 python tools/prof_driver.py 256 3 struct 2 2>&1 | grep " rep " | sort -t: -k2 -n | head -1
