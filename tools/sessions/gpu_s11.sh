#!/bin/bash
# the share of one GPU in the strong-scaled 512^3 step on 8 GPUs (8 boxes of 128^3): kernel span and drain tail against the chunk size of the work queue
mkdir -p gpurun_out
{
for ch in 256 64 32; do
  echo "== NYX_HC_CHUNK=$ch"
  NYX_HC_CHUNK=$ch timeout 600 python bench.py --n 256 --box 128 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for p,r in d['paths'].items(): print(p, 'ms_per_step %.3f' % r['ms_per_step'], 'kernel_ms %.3f' % r['kernel_ms_this_rank'], 'drain_tail_ms %.3f' % r['drain_tail_ms_this_rank'], 'value %.4g' % r['value'])
"
done
echo "== 512^3, chunk 256 / 64"
for ch in 256 64; do
NYX_HC_CHUNK=$ch timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for p,r in d['paths'].items(): print('$ch', p, 'ms_per_step %.3f' % r['ms_per_step'], 'kernel_ms %.3f' % r['kernel_ms_this_rank'], 'drain_tail_ms %.3f' % r['drain_tail_ms_this_rank'], 'value %.4g' % r['value'])
"
done
} > gpurun_out/s11.log 2>&1
cat gpurun_out/s11.log
