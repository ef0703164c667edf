#!/bin/bash
# session 19: which warps end the bookkeeping phase (first / last lane class of the slowest warp per CTA-round)
mkdir -p gpurun_out
{
  f=build/variants/libnyx_hc_timingM.so
  echo "== vec"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 vec 3 2>&1 | tail -75
  echo "== struct"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 struct 3 2>&1 | tail -75
} > gpurun_out/s19.log 2>&1
tail -5 gpurun_out/s19.log
