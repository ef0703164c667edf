#!/bin/bash
# bench.py as the driver runs it (N=1), the reference arm, ncu launch list + full captures at the bench's size (traffic)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s7_bench_ref.json 2> gpurun_out/s7_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("s7_bench", "s7_bench_ref"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.4g" % d["value"], "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}))
        for p, r in (d.get("paths") or {}).items():
            print("   path", p, "value %.4g" % r["value"], "frac %.4f" % r["roofline"]["frac"], "e2e", (r.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_512cubed.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/s7_bench_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hc_sorted_kernelILi0 -s 1 -c 1 -f -o gpurun_out/r2_sorted_vec_512 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --paths one > gpurun_out/s7_ncu_vec.log 2>&1; tail -1 gpurun_out/s7_ncu_vec.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hc_sorted_kernelILi1 -s 1 -c 1 -f -o gpurun_out/r2_sorted_struct_512 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --paths one --path struct > gpurun_out/s7_ncu_struct.log 2>&1; tail -1 gpurun_out/s7_ncu_struct.log
