#!/bin/bash
# session 18: compute-sanitizer memcheck over the kernels added this round (SAVE_REACT instantiation + assembly kernel, conservative
# enforce_minimum_density, split source sweeps), then the final ncu captures of the two integrator kernels at the bench's size
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conservative or save_react or sources_bitwise" > gpurun_out/s18_memcheck.log 2>&1; echo "memcheck rc=$?" > gpurun_out/s18.log
tail -4 gpurun_out/s18_memcheck.log >> gpurun_out/s18.log
for p in vec struct; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -c 1 -o gpurun_out/r2b_sorted_${p}_512 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --paths one --path $p > gpurun_out/s18_ncu_$p.log 2>&1; echo "ncu $p rc=$?" >> gpurun_out/s18.log
done
ls -la gpurun_out/*.ncu-rep >> gpurun_out/s18.log 2>&1
cat gpurun_out/s18.log
