#!/bin/bash
# MODE_MN_TMA: per-layer times by level, then tests + bench at level 1
mkdir -p gpurun_out
for lvl in 0 2; do
  echo "== PZ_TMA_FPROP=$lvl"
  for l in 8 9 13 14; do PZ_TMA_FPROP=$lvl timeout 120 python tools/bench_layers.py 64 $l 2>&1 | grep -v "^layer\|^sum"; done
done > gpurun_out/r4d_layers.txt 2>&1
cat gpurun_out/r4d_layers.txt
PZ_TMA_FPROP=2 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_parity_cuda.py tests/test_gpu_nets.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r4d_pytest.txt
for lvl in 0 1 2; do
PZ_TMA_FPROP=$lvl timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print('level $lvl', d['value'], d['ms_per_step'], d['e2e']['value'], f)"
done 2>&1 | tee gpurun_out/r4d_bench.txt
true
