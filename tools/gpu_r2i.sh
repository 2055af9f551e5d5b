#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
head -c 500 gpurun_out/r2i_bench_2gpu.json; tail -n 5 gpurun_out/r2i_bench_2gpu.err
true
