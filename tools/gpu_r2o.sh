#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "conv or parity or nets or gemm or deconv" 2>&1 | tail -15 > gpurun_out/r2o_pytest.log
PZ_NO_VEC_GATHER=1 timeout 600 python tools/bench_layers.py > gpurun_out/r2o_layers_novec.txt 2>&1
timeout 600 python tools/bench_layers.py > gpurun_out/r2o_layers_vec.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
tail -n 4 gpurun_out/r2o_pytest.log; tail -n 3 gpurun_out/r2o_layers_novec.txt; tail -n 3 gpurun_out/r2o_layers_vec.txt; head -c 300 gpurun_out/r2o_bench.json
true
