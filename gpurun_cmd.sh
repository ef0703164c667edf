set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for v in t512L2 t384L2 t256L2 t384L t384 t512; do
  NYX_HC_LIB=$PWD/build/variants/libnyx_hc_$v.so timeout 120 python bench.py --n 256 --box 64 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
  python -c "import json;d=json.load(open('gpurun_out/var_$v.json'));print('$v', d['value'], d['ms_per_step'], d['roofline']['frac'], d['stats']['sum_nst'])" || tail -3 gpurun_out/var_$v.err
done
timeout 300 python tests/gpu_report.py 2>&1 | tail -3
cp build/variants/libnyx_hc_t512L2.so gpurun_out/libnyx_hc_profiled.so
NYX_HC_LIB=$PWD/build/variants/libnyx_hc_t512L2.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:hc_integrate -s 1 -c 1 -o gpurun_out/prof_r1f python bench.py --n 128 --box 64 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
