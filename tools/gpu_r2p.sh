#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "conv or parity or nets or gemm or deconv" 2>&1 | tail -15 > gpurun_out/r2p_pytest.log
timeout 600 python tools/bench_layers.py > gpurun_out/r2p_layers_default.txt 2>&1
PZ_NO_STAGED_EPI=1 timeout 600 python tools/bench_layers.py > gpurun_out/r2p_layers_nostaged.txt 2>&1
PZ_NO_VEC_GATHER=1 timeout 600 python tools/bench_layers.py > gpurun_out/r2p_layers_novec.txt 2>&1
for l in 3 4 8 9 13 14 18; do
  PZB200_LIB=$PWD/puzzlelib_b200/libpzb200_timeline.so PZ_NO_VEC_GATHER=1 timeout 300 python tools/bench_layers.py 64 $l >> gpurun_out/r2p_timeline.txt 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
tail -n 4 gpurun_out/r2p_pytest.log; for f in default nostaged novec; do tail -n 1 gpurun_out/r2p_layers_$f.txt; done; head -c 300 gpurun_out/r2p_bench.json
true
