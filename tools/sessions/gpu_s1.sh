#!/bin/bash
# round-2 session 1: baseline + phase/stage timing builds + -fmad=true variant (speed and parity)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
nproc > gpurun_out/s1_nproc.txt; lscpu | head -20 >> gpurun_out/s1_nproc.txt
{
for v in default timing2 timing0 fmad; do
  if [ $v = default ]; then f=nyx_b200/csrc/libnyx_hc.so; else f=build/variants/libnyx_hc_$v.so; fi
  echo "== $v"
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 256 4 vec 3 2>&1
  HC_LIB=$PWD/$f timeout 300 python tools/prof_driver.py 128 4 struct 3 2>&1
done
echo "== parity report, default"
timeout 600 python tests/gpu_report.py 2>&1
echo "== parity report, fmad=true"
NYX_HC_LIB=$PWD/build/variants/libnyx_hc_fmad.so timeout 600 python tests/gpu_report.py 2>&1
} > gpurun_out/s1.log 2>&1
tail -5 gpurun_out/s1.log
