#!/bin/bash
# GPU session: smoke, full parity suite, default bench (512^3), SDC bench, rank-2/4 rows, redshift sweep (config 5)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_sources.py 128 16 7 > gpurun_out/sources_rows.json 2> gpurun_out/sources_rows.err; echo "bench_sources rc=$?"; cut -c1-700 gpurun_out/sources_rows.json
timeout 600 python tools/redshift_sweep.py 256 128 3 > gpurun_out/redshift_sweep.json 2> gpurun_out/redshift_sweep.err; echo "sweep rc=$?"; cut -c1-200 gpurun_out/redshift_sweep.json
timeout 600 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_512.json
timeout 600 python bench.py --path struct --no-cpu --steps 3 --warmup 3 > gpurun_out/bench_512_struct.json 2> gpurun_out/bench_512_struct.err; echo "struct rc=$?"; cut -c1-300 gpurun_out/bench_512_struct.json
