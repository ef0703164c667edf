#!/bin/bash
# session 26: GPU suite + quick timing after the source clean-up (codegen unchanged)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s26_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s26_pytest_gpu.log
timeout 300 python tools/prof_driver.py 256 5 vec 3 2>&1 | grep " rep " | tail -2
timeout 300 python tools/prof_driver.py 256 5 struct 3 2>&1 | grep " rep " | tail -2
