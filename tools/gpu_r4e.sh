#!/bin/bash
# 16-bit wgrad operands through the copy engine: 16-bit tests, side benches with / without; then the full GPU tier, smoke, headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | tee gpurun_out/r4e_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
for lvl in 0 2; do
for cfg in "vgg16 bf16 128" "resnet50 bf16 64"; do set -- $cfg
PZ_TMA_WGRAD=$lvl timeout 600 python bench.py --model $1 --dtype $2 --batch $3 --steps 10 --warmup 3 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PZ_TMA_WGRAD=$lvl $1 $2', d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done 2>&1 | tee gpurun_out/r4e_side.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4e_bench.json 2> gpurun_out/r4e_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r4e_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
true
