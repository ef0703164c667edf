#!/bin/bash
mkdir -p gpurun_out
{
for v in pre8 nopf default; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep " | tail -3
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 3 2>&1 | grep " rep " | tail -2
done
} > gpurun_out/s8.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/r2_sorted_vec_512 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --paths one > gpurun_out/s8_ncu_vec.log 2>&1; tail -1 gpurun_out/s8_ncu_vec.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/r2_sorted_struct_512 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --paths one --path struct > gpurun_out/s8_ncu_struct.log 2>&1; tail -1 gpurun_out/s8_ncu_struct.log
cat gpurun_out/s8.log
