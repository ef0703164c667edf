#!/bin/bash
# GPU session: parity tests, bench at 512^3, ncu launch list, ncu --set full of the dominant kernel (128^3)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_512.json
timeout 300 python tools/prof_driver.py 256 3 vec 3 > gpurun_out/drv_256.log 2>&1; cat gpurun_out/drv_256.log
timeout 300 python tools/prof_driver.py 128 3 struct 3 > gpurun_out/drv_128s.log 2>&1; cat gpurun_out/drv_128s.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_256.csv python bench.py --n 256 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/prof_sorted_128 python tools/prof_driver.py 128 2 vec 3 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
