#!/bin/bash
# round 2, re-entry: copy-engine (3-d tensor map) operands for wgrad -- correctness, per-layer times, whole-step effect
mkdir -p gpurun_out
for lvl in 0 1 2; do
  PZ_TMA_WGRAD=$lvl timeout 120 python tools/check_tma_wgrad.py 2>&1 | tail -12
done > gpurun_out/r4a_check.txt 2>&1
cat gpurun_out/r4a_check.txt
if grep -q "FAIL\|Error\|error" gpurun_out/r4a_check.txt || [ $(grep -c "OK" gpurun_out/r4a_check.txt) -lt 3 ]; then echo "CHECK FAILED"; exit 0; fi
for lvl in 0 1 2; do
  echo "== PZ_TMA_WGRAD=$lvl"
  for l in 7 8 9 12 13 14; do PZ_TMA_WGRAD=$lvl timeout 120 python tools/bench_layers.py 64 $l 2>&1 | grep -v "^layer\|^sum"; done
done > gpurun_out/r4a_layers.txt 2>&1
cat gpurun_out/r4a_layers.txt
PZ_TMA_WGRAD=2 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_parity_cuda.py -m gpu -q -x -k "conv or wgrad or resnet or layer" 2>&1 | tail -5 | tee gpurun_out/r4a_pytest.txt
for lvl in 0 2; do
PZ_TMA_WGRAD=$lvl timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print('level $lvl', d['value'], d['ms_per_step'], d['e2e']['value'], f)"
done 2>&1 | tee gpurun_out/r4a_bench.txt
true
