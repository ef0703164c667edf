#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ab_ring_$i.json 2> gpurun_out/ab_ring_$i.err
  NYX_HC_PAGEABLE_DESC=1 timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ab_pageable_$i.json 2> gpurun_out/ab_pageable_$i.err
done
python - <<'PY'
import json
for n in ("ring_1", "pageable_1", "ring_2", "pageable_2"):
    d = json.load(open(f"gpurun_out/ab_{n}.json"))
    print(n, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "e2e ms", round(d["e2e"]["ms_per_step"], 1))
PY
