#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err
PZ_GRID_NO_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_2gpu_nooverlap.json 2> gpurun_out/r2j_bench_2gpu_nooverlap.err
head -c 300 gpurun_out/r2j_bench_2gpu.json; echo; head -c 300 gpurun_out/r2j_bench_2gpu_nooverlap.json; echo; tail -n 4 gpurun_out/r2j_bench_2gpu.err
true
