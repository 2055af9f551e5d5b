#!/bin/bash
# round 2, call C: the whole GPU test tier on the reference tree over the seam, the reference's unit tests, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_seam_reference_unittests.py 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
timeout 1200 python -m pytest tests/test_gpu_seam_reference_unittests.py -q 2>&1 | tail -80 > gpurun_out/r2c_pytest_seam.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -n 5 gpurun_out/r2c_pytest.log gpurun_out/r2c_pytest_seam.log gpurun_out/r2c_smoke.log; head -c 600 gpurun_out/r2c_bench.json; tail -n 5 gpurun_out/r2c_bench.err
true
