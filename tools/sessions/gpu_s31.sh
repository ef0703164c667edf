#!/bin/bash
# session 31: host pipeline stages only the tiles' bounding box of a ghosted FAB: full GPU suite
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s31_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s31_pytest.log
