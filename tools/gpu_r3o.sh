#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "deferred or pending or folded or nets or parity or Add or Replicate or ResNet or activations or batchnorm or BatchNorm or CuDnnNorm or instancenorm or InstanceNorm or fullsize" 2>&1 | tail -8
for cfg in "PZ_NO_SUM_RELU_FUSION=1" "PZ_NO_SUM_RELU_FUSION=0" "PZ_NO_SUM_RELU_FUSION=1" "PZ_NO_SUM_RELU_FUSION=0"; do
env $cfg timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print('$cfg', round(d['value'],1), round(d['ms_per_step'],3), 'eltwise', round(f['eltwise'],3), d['gpu_launches'])"
done
true
