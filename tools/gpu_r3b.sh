#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r3b_pytest.log
tail -n 30 gpurun_out/r3b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
true
