#!/bin/bash
# GPU session: parity tests, bench at 512^3 (N=1), ncu launch list of the bench command, ncu --set full of the dominant kernel (256^3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_512.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_512.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hc_sorted -s 1 -c 1 -f -o gpurun_out/prof_sorted_256 python tools/prof_driver.py 256 2 vec 3 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
