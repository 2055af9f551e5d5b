#!/bin/bash
# copied fp32 tiles rounded in place (FIXUP): error checks, full GPU tier, smoke, headline bench
mkdir -p gpurun_out
{ PZ_TMA_WGRAD=2 timeout 120 python tools/check_tma_wgrad.py f32 2>&1 | tail -13; PZ_TMA_FPROP=2 timeout 120 python tools/check_tma_fprop.py 2>&1 | tail -9; } | tee gpurun_out/r4k_check.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 | tee gpurun_out/r4k_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4k_bench.json 2> gpurun_out/r4k_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r4k_bench.json').read().strip().splitlines()[-1]);print('final', d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['roofline']['frac'], d['roofline']['families_ms_per_step'])"
true
