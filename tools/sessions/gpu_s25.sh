#!/bin/bash
# session 25: compute-sanitizer racecheck + synccheck of the two integrator kernels (shared-memory lane state, two barriers per round)
mkdir -p gpurun_out
{
for tool in racecheck synccheck; do
  for p in vec struct; do
    echo "== $tool $p"
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 python tools/prof_driver.py 48 1 $p 3 2>&1 | grep -v "^$" | tail -6
    echo "rc=$?"
  done
done
} > gpurun_out/s25.log 2>&1
cat gpurun_out/s25.log
