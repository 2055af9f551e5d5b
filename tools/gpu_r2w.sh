#!/bin/bash
# N-GPU bench exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2w_bench_${N}gpu.json 2> gpurun_out/r2w_bench_${N}gpu.err
tail -5 gpurun_out/r2w_bench_${N}gpu.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2w_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('gradient_sync'), d['config'].get('param_checksum_equal_across_ranks'))"
true
