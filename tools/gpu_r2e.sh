#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150 > gpurun_out/r2e_pytest.log
timeout 600 python tools/bench_ops.py > gpurun_out/r2e_ops.log 2>&1
PZ_BN_NO_CLUSTER=1 timeout 600 python tools/bench_ops.py 64 bn > gpurun_out/r2e_ops_old_bn.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -n 4 gpurun_out/r2e_pytest.log; tail -n 2 gpurun_out/r2e_ops.log gpurun_out/r2e_ops_old_bn.log; head -c 400 gpurun_out/r2e_bench.json; tail -n 3 gpurun_out/r2e_bench.err
true
