#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 100 0" "16 50 0" "16 50 256" "8 100 256" "16 66 0"; do
  set -- $cfg
  echo "== MAX_CL=$1 STASH_KB=$2 THREADS=$3" >> gpurun_out/r2m_bn.log
  PZ_BN_MAX_CL=$1 PZ_BN_STASH_KB=$2 PZ_BN_THREADS=$3 timeout 300 python bench.py --steps 10 --warmup 3 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']['families_ms_per_step']; print(d['ms_per_step'], 'bn_fwd', r['bn_fwd'], 'bn_bwd', r['bn_bwd'])" >> gpurun_out/r2m_bn.log 2>&1
done
cat gpurun_out/r2m_bn.log
