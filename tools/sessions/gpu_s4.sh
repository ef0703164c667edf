#!/bin/bash
mkdir -p gpurun_out
{
echo "== default (struct fix)"
timeout 300 python tools/prof_driver.py 256 4 vec 3 2>&1 | grep " rep "
timeout 300 python tools/prof_driver.py 128 4 struct 3 2>&1 | grep " rep "
} > gpurun_out/s4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/r2a_sorted_vec_256 python tools/prof_driver.py 256 2 vec 3 > gpurun_out/s4_ncu_vec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/r2a_sorted_struct_128 python tools/prof_driver.py 128 2 struct 3 > gpurun_out/s4_ncu_struct.log 2>&1
cat gpurun_out/s4.log; tail -2 gpurun_out/s4_ncu_vec.log
