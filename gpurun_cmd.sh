timeout 120 python tools/gpu_diag_eos.py 2>&1 | tail -12
