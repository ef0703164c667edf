#!/bin/bash
# session 15: conservative enforce_minimum_density (kernel + real-executable run), SAVE_REACT retest, then the full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "conservative or sources_dropin or react" > gpurun_out/s15_new.log 2>&1; echo "new rc=$?" > gpurun_out/s15.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s15.log
cat gpurun_out/s15.log; tail -30 gpurun_out/s15_new.log; tail -5 gpurun_out/s15_pytest.log
