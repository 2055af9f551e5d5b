#!/bin/bash
mkdir -p gpurun_out
timeout 50 python bench.py --model vgg16 --dtype f16 --batch 128 --steps 10 --warmup 3 --no-ref-gpu --no-cpu > gpurun_out/r4n_vgg16_f16.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r4n_vgg16_f16.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['dtype'][:8])"
true
