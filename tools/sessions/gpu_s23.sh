#!/bin/bash
# session 23: host pipeline with geometric ramps at both ends (HC_HOST_TAPER=2) against the default, N = 1
mkdir -p gpurun_out
for i in 1 2; do
  timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/s23_def_$i.json 2> gpurun_out/s23_def_$i.err
  NYX_HC_LIB=$PWD/build/variants/libnyx_hc_taper2.so timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/s23_tap_$i.json 2> gpurun_out/s23_tap_$i.err
done
python - > gpurun_out/s23.log <<'PY'
import json
for n in ("def_1", "tap_1", "def_2", "tap_2"):
    try:
        d = json.load(open(f"gpurun_out/s23_{n}.json"))
        s = d["paths"]["struct"]
        print(n, "vec value %.4g e2e %.4g (%.1f ms); struct value %.4g e2e %.4g (%.1f ms)" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], s["value"], s["e2e"]["value"], s["e2e"]["ms_per_step"]))
    except Exception as e:
        print(n, "failed", e)
PY
cat gpurun_out/s23.log
