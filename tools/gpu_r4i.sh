#!/bin/bash
mkdir -p gpurun_out
T=tools/ubench/tma4d
{
echo "# tools/ubench/tma4d: one 4-d tiled TMA load of a float tensor (W H C N | box w h c | coordinates | swizzle 0 none / 3 128B)"
$T 28 28 128 8  32 1 128  0 0 0 0  3
$T 28 28 128 8  32 1 128  -1 -1 0 0  3
$T 28 28 128 8  32 1 128  1 0 0 0  3
$T 28 28 128 8  32 1 128  4 0 0 0  3
$T 28 28 128 8  32 1 128  0 -1 0 0  3
$T 28 28 128 8  32 1 128  4 -1 0 0  3
$T 28 28 128 8  32 1 128  -4 0 0 0  3
$T 28 28 128 8  32 1 128  -4 -2 0 0  0
$T 28 28 128 8  8 4 128  0 0 0 0  3
$T 28 28 128 8  8 4 128  0 0 0 0  0
$T 64 28 128 8  32 1 128  40 5 0 1  3
} > gpurun_out/r4i_tma4d.txt 2>&1
cat gpurun_out/r4i_tma4d.txt
timeout 600 python -m pytest tests/test_gpu_tma_operands.py -m gpu -q -x 2>&1 | tail -5
for lvl in 0 1 2; do
PZ_TMA_FPROP=$lvl timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print('PZ_TMA_FPROP=$lvl', d['value'], d['ms_per_step'], d['e2e']['value'], f)"
done 2>&1 | tee gpurun_out/r4i_bench.txt
true
