#!/bin/bash
# wgrad of stride-1 multi-tap filters with both operands as patches through 4-d tensor maps (PZ_TMA_WGRAD=3)
mkdir -p gpurun_out
{
for lvl in 2 3; do for dt in f32 f16; do
PZ_TMA_WGRAD=$lvl timeout 120 python tools/check_tma_wgrad.py $dt 2>&1 | tail -14
done; done
} > gpurun_out/r4f_check.txt 2>&1
cat gpurun_out/r4f_check.txt
if grep -q "FAIL\|Error\|error" gpurun_out/r4f_check.txt; then echo "CHECK FAILED"; exit 0; fi
for lvl in 2 3; do
  echo "== PZ_TMA_WGRAD=$lvl"
  PZ_TMA_WGRAD=$lvl timeout 120 python tools/bench_layers.py 64 7 2>&1 | grep -v "^layer\|^sum"
done 2>&1 | tee gpurun_out/r4f_layers.txt
PZ_TMA_WGRAD=3 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_parity_cuda.py tests/test_gpu_nets.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r4f_pytest.txt
for lvl in 2 3; do
for cfg in "vgg16 bf16 128" "resnet50 f32 64"; do set -- $cfg
PZ_TMA_WGRAD=$lvl timeout 600 python bench.py --model $1 --dtype $2 --batch $3 --steps 10 --warmup 3 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PZ_TMA_WGRAD=$lvl $1 $2', d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done 2>&1 | tee gpurun_out/r4f_side.txt
true
