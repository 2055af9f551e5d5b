#!/bin/bash
# full validation of the round-2 state: GPU tier, smoke, both bench arms, the reference's unit tests over the seam
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r3z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r3z_smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3z_bench_reference.json 2> gpurun_out/r3z_bench_reference.err
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err
timeout 1500 python tools/run_ref_unittests.py --impl b200 --out gpurun_out/r3z_ref_unittests_b200.json > gpurun_out/r3z_ref_unittests.log 2>&1
tail -n 6 gpurun_out/r3z_pytest.log; tail -n 2 gpurun_out/r3z_smoke.log; head -c 400 gpurun_out/r3z_bench.json; echo; head -c 300 gpurun_out/r3z_bench_reference.json; echo; tail -n 3 gpurun_out/r3z_ref_unittests.log
true
