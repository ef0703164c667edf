mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sources or fab_copy or multi_rank or host_buffer" > gpurun_out/pytest_rank2.log 2>&1; echo "pytest rank2 rc=$?"; tail -15 gpurun_out/pytest_rank2.log
timeout 300 python tools/bench_sources.py 128 16 7 > gpurun_out/sources_rows.json 2> gpurun_out/sources_rows.err; echo "bench_sources rc=$?"; cat gpurun_out/sources_rows.json; tail -3 gpurun_out/sources_rows.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_sources -s 2 -c 2 -f -o gpurun_out/prof_sources python tools/bench_sources.py 128 16 1 --no-host > gpurun_out/ncu_sources.log 2>&1; tail -2 gpurun_out/ncu_sources.log
