#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "batchnorm or bn or parity or nets or instancenorm or BatchNorm or CuDnnNorm or fullsize" 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v,3) for k,v in f.items()})"
true
