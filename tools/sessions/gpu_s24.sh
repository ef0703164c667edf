#!/bin/bash
# session 24: final check of the tree as committed: smoke, GPU suite, bench with defaults (as the driver runs it)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s24_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s24_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -s -rA > gpurun_out/s24_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s24_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/s24_bench_512.json 2> gpurun_out/s24_bench_512.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/s24_bench_512.json"))
print("value %.4g" % d["value"], "ms", d.get("ms_per_step"), "e2e %.4g" % d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "clocks", d.get("clocks"))
for p, r in (d.get("paths") or {}).items():
    print("   path", p, "value %.4g" % r["value"], "frac %.4f" % r["roofline"]["frac"], "e2e %.4g" % r["e2e"]["value"], r["e2e"]["ms_per_step"])
PY
