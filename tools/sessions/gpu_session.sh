#!/bin/bash
# GPU session of a round: smoke(), parity tests, bench at 512^3 (N=1), reference arm, SDC-path bench, the rows either side of the path
# (SURVEY 8f ranks 1, 2), the config-5 redshift sweep, ncu launch list of the bench command, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --path struct --no-cpu --steps 3 --warmup 3 > gpurun_out/bench_512_struct.json 2> gpurun_out/bench_512_struct.err; echo "struct rc=$?"
python - <<'PY'
import json
for f in ("bench_512", "bench_512_struct", "bench_ref"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d["value"], d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 python tools/bench_eos_rows.py 128 16 5 > gpurun_out/eos_rows.json 2> gpurun_out/eos_rows.err; echo "eos rows rc=$?"
timeout 300 python tools/bench_sources.py 128 16 7 > gpurun_out/sources_rows.json 2> gpurun_out/sources_rows.err; echo "sources rows rc=$?"
timeout 600 python tools/redshift_sweep.py 256 128 3 > gpurun_out/redshift_sweep.json 2> gpurun_out/redshift_sweep.err; echo "sweep rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_512.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/prof_sorted_256 python tools/prof_driver.py 256 2 vec 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/prof_sorted_struct_128 python tools/prof_driver.py 128 2 struct 3 > gpurun_out/ncu_full_struct.log 2>&1; tail -2 gpurun_out/ncu_full_struct.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_sources -s 2 -c 1 -f -o gpurun_out/prof_sources python tools/bench_sources.py 128 16 1 --no-host > gpurun_out/ncu_sources.log 2>&1; tail -1 gpurun_out/ncu_sources.log
