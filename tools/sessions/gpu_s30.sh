#!/bin/bash
# session 30: the real-executable tests after the rebuild
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dropin_real.py -m gpu -q > gpurun_out/s30_real.log 2>&1; echo "real rc=$?"; tail -5 gpurun_out/s30_real.log
