#!/bin/bash
# full GPU tier + smoke + bench with the copy-engine wgrad operands on by default
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 | tee gpurun_out/r4b_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4b_bench.json 2> gpurun_out/r4b_bench.err
tail -c 3000 gpurun_out/r4b_bench.json
true
