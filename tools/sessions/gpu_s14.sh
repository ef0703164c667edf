#!/bin/bash
# session 14: SAVE_REACT entry points (new tests first), then the full GPU suite; production kernels re-timed (must be unchanged)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "react" > gpurun_out/s14_react.log 2>&1; echo "react rc=$?" > gpurun_out/s14.log
{
  timeout 300 python tools/prof_driver.py 256 6 vec 3 2>&1 | grep " rep " | tail -3
  timeout 300 python tools/prof_driver.py 256 5 struct 3 2>&1 | grep " rep " | tail -3
} >> gpurun_out/s14.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s14.log
cat gpurun_out/s14.log; tail -30 gpurun_out/s14_react.log; tail -5 gpurun_out/s14_pytest.log
