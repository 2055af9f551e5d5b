#!/bin/bash
mkdir -p gpurun_out
for cfg in "PZ_BN_STREAM_STORES=0" "PZ_BN_STREAM_STORES=1" "PZ_BN_STREAM_STORES=0" "PZ_BN_STREAM_STORES=1"; do
env $cfg timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print('$cfg', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in f.items()})"
done
true
