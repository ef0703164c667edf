#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("value %.4g" % d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"))
for p, r in (d.get("paths") or {}).items():
    print("   path", p, "value %.4g" % r["value"], "frac %.4f" % r["roofline"]["frac"], "e2e", (r.get("e2e") or {}))
print("strong", json.dumps(d.get("strong"), indent=1))
PY
tail -5 gpurun_out/bench_n$N.err
