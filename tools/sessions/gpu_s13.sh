#!/bin/bash
# session 13: stage-boundary variants, host pipeline with alternating compute streams (A/B against one stream, 8 vs 16 groups), full GPU test suite
mkdir -p gpurun_out
{
for v in default sync0 sync1; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 6 vec 3 2>&1 | grep " rep " | tail -3
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 5 struct 3 2>&1 | grep " rep " | tail -3
done
} > gpurun_out/s13.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13.log
for i in 1 2; do
  timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/s13_two_$i.json 2> gpurun_out/s13_two_$i.err
  NYX_HC_HOST_ONE_COMP_STREAM=1 timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/s13_one_$i.json 2> gpurun_out/s13_one_$i.err
  NYX_HC_LIB=$PWD/build/variants/libnyx_hc_groups16.so timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/s13_g16_$i.json 2> gpurun_out/s13_g16_$i.err
done
python - >> gpurun_out/s13.log <<'PY'
import json
for n in ("two_1", "one_1", "g16_1", "two_2", "one_2", "g16_2"):
    try:
        d = json.load(open(f"gpurun_out/s13_{n}.json"))
        s = d["paths"]["struct"]
        print(n, "vec value %.4g e2e %.4g (%.1f ms); struct value %.4g e2e %.4g (%.1f ms)" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], s["value"], s["e2e"]["value"], s["e2e"]["ms_per_step"]))
    except Exception as e:
        print(n, "failed", e)
PY
cat gpurun_out/s13.log; tail -3 gpurun_out/s13_pytest.log
