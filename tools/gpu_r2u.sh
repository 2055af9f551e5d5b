#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2u_bn.txt
for mb in 24 48 80 110; do
  echo "== PZ_BN_SLAB_MB=$mb" >> gpurun_out/r2u_bn.txt
  PZ_BN_SLAB_MB=$mb timeout 600 python tools/bench_ops.py 64 bn 2>&1 | grep "112x112\|55x55\|totals" >> gpurun_out/r2u_bn.txt
  PZ_BN_SLAB_MB=$mb timeout 900 python bench.py --steps 10 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print(d['ms_per_step'], 'bn_fwd', f['bn_fwd'], 'bn_bwd', f['bn_bwd'])" >> gpurun_out/r2u_bn.txt
done
cat gpurun_out/r2u_bn.txt
true
