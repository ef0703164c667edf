#!/bin/bash
mkdir -p gpurun_out
HC_LIB=$PWD/build/variants/libnyx_hc_a384.so timeout 600 python tools/redshift_sweep.py 256 128 3 > gpurun_out/redshift_sweep_a384.json 2> gpurun_out/redshift_sweep_a384.err; echo "sweep a384 rc=$?"
timeout 600 python tools/redshift_sweep.py 256 128 3 > gpurun_out/redshift_sweep_a320.json 2> gpurun_out/redshift_sweep_a320.err; echo "sweep a320 rc=$?"
python - <<'PY'
import json
for v in ("a320", "a384"):
    rows = [json.loads(l) for l in open(f"gpurun_out/redshift_sweep_{v}.json")]
    print(v, " ".join(f"{r['z']}:{r['ms_median']:.1f}" for r in rows), "sum", round(sum(r["ms_median"] for r in rows), 1))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hc_sources|hc_fab_op" -c 60 --csv --log-file gpurun_out/launches_sources.csv python tools/bench_sources.py 128 16 1 --no-host > gpurun_out/ncu_sources_list.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_sources.csv")) if len(r) > 5 and r[0].isdigit()]
import collections
d = collections.defaultdict(list)
for r in rows:
    d[r[4][:60]].append(float(r[-1]))
for k, v in d.items():
    print(k, len(v), "median", sorted(v)[len(v)//2], "min", min(v), "max", max(v))
PY
