#!/bin/bash
# session 28: compute-sanitizer initcheck (reads of uninitialised device memory) over the integrator, SAVE_REACT and conservative-variant tests
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conservative or save_react or struct_matches or vec_matches or eos_kernel" > gpurun_out/s28_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -c "Uninitialized" gpurun_out/s28_initcheck.log; grep -A12 "Uninitialized" gpurun_out/s28_initcheck.log | head -60; tail -4 gpurun_out/s28_initcheck.log
