#!/bin/bash
# round-2 final single-GPU session: smoke, the full GPU test suite (log), bench.py as the driver runs it, the reference arm, launch list, config-5 sweep
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f1_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f1_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -s -rA > gpurun_out/f1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f1_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/f1_bench_512.json 2> gpurun_out/f1_bench_512.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/f1_bench_ref.json 2> gpurun_out/f1_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 --path struct > gpurun_out/f1_bench_ref_struct.json 2> gpurun_out/f1_bench_ref_struct.err; echo "ref struct rc=$?"
python - <<'PY'
import json
for f in ("f1_bench_512", "f1_bench_ref", "f1_bench_ref_struct"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.4g" % d["value"], "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}))
        for p, r in (d.get("paths") or {}).items():
            print("   path", p, "value %.4g" % r["value"], "frac %.4f" % r["roofline"]["frac"], "e2e", (r.get("e2e") or {}).get("value"), "drain", r.get("drain_tail_ms_this_rank"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_512cubed.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/f1_bench_ncu.log 2>&1
timeout 600 python tools/redshift_sweep.py 256 128 3 > gpurun_out/f1_redshift_sweep.json 2> gpurun_out/f1_redshift_sweep.err; echo "sweep rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/f1_smi.txt 2>&1
