#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ghosted or host_buffer" > gpurun_out/s32.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/s32.log
