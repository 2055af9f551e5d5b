#!/bin/bash
# side configurations on the final code (refreshes profiles/r02_bench_{vgg16,resnet50}_bf16.json)
mkdir -p gpurun_out
timeout 55 python bench.py --model vgg16 --dtype bf16 --batch 128 --steps 10 --warmup 3 --no-ref-gpu --no-cpu > gpurun_out/r4m_vgg16_bf16.json 2>/dev/null
timeout 45 python bench.py --model resnet50 --dtype bf16 --batch 64 --steps 10 --warmup 3 --no-ref-gpu --no-cpu > gpurun_out/r4m_resnet50_bf16.json 2>/dev/null
for f in gpurun_out/r4m_vgg16_bf16.json gpurun_out/r4m_resnet50_bf16.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['dtype'][:8])"; done
true
