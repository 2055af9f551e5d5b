#!/bin/bash
# BASELINE.json configs[3]: ResNet-50 in 16-bit storage, data parallel over 8 GPUs
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 8 --steps 20 --warmup 5 --dtype bf16 --no-cpu > gpurun_out/r3k_bench_8gpu_bf16.json 2> gpurun_out/r3k_bench_8gpu_bf16.err
tail -3 gpurun_out/r3k_bench_8gpu_bf16.err
python -c "
import json
d=json.loads(open('gpurun_out/r3k_bench_8gpu_bf16.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['dtype'][:20], d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('param_checksum_equal_across_ranks'))"
true
