#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "batchnorm or bn or parity or nets or instancenorm" 2>&1 | tail -30 > gpurun_out/r2l_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:cluster --csv --log-file gpurun_out/r2l_bn_ncu.csv python tools/bench_ops.py 64 bn > /dev/null 2>&1
tail -n 4 gpurun_out/r2l_pytest.log; head -c 300 gpurun_out/r2l_bench.json; tail -n 3 gpurun_out/r2l_bench.err
true
