#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -80 > gpurun_out/r2g_pytest.log
timeout 300 python tools/bench_ops.py 64 bn > gpurun_out/r2g_ops_bn.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
PZ_BN_NO_CLUSTER=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu > gpurun_out/r2g_bench_oldbn.json 2> gpurun_out/r2g_bench_oldbn.err
tail -n 8 gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_ops_bn.log | tail -20; head -c 300 gpurun_out/r2g_bench.json; tail -n 3 gpurun_out/r2g_bench.err
true
