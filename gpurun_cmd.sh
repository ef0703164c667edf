set -x
mkdir -p gpurun_out
cp nyx_b200/csrc/libnyx_hc.so gpurun_out/libnyx_hc_profiled.so
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 300 python tools/gpu_diag_flags.py 2>&1 | tail -24
timeout 1200 python bench.py > gpurun_out/bench_512_default.json 2> gpurun_out/bench_512.err; tail -3 gpurun_out/bench_512.err; cat gpurun_out/bench_512_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --n 256 --box 64 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_integrate -s 1 -c 1 -o gpurun_out/prof_r1_default python bench.py --n 128 --box 64 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
