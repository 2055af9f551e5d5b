#!/bin/bash
# batch-norm backward with absorbed parameter-gradient accumulation: full GPU tier, smoke, A/B bench, headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 | tee gpurun_out/r4j_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
PZ_NO_BN_ACC_FUSION=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print('PZ_NO_BN_ACC_FUSION=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], f)" | tee gpurun_out/r4j_ab.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4j_bench.json 2> gpurun_out/r4j_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r4j_bench.json').read().strip().splitlines()[-1]);print('fused', d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline']['frac'], d['roofline']['families_ms_per_step'])" | tee -a gpurun_out/r4j_ab.txt
true
