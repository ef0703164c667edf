#!/bin/bash
# session 29: the AMREX_USE_GPU flavour of the drop-in with device FABs, then the drop-in tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py -m gpu -q -x > gpurun_out/s29_dropin.log 2>&1; echo "dropin rc=$?"; tail -15 gpurun_out/s29_dropin.log
