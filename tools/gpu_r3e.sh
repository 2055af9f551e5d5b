#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "batchnorm or bn or parity or nets or instancenorm or BatchNorm or CuDnnNorm or fullsize" 2>&1 | tail -15 > gpurun_out/r3e_pytest.log
timeout 600 python tools/bench_ops.py 64 bn > gpurun_out/r3e_bn_persistent.txt 2>&1
PZ_BN_PERSISTENT=0 timeout 600 python tools/bench_ops.py 64 bn > gpurun_out/r3e_bn_oneshot.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err
tail -n 6 gpurun_out/r3e_pytest.log; paste gpurun_out/r3e_bn_persistent.txt gpurun_out/r3e_bn_oneshot.txt | tail -20; head -c 300 gpurun_out/r3e_bench.json; tail -3 gpurun_out/r3e_bench.err
true
