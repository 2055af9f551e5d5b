#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
rm -f gpurun_out/r2x_sweep.txt
run() {
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config'].get('param_checksum_equal_across_ranks'))" >> gpurun_out/r2x_sweep.txt
}
run PZ_DUMMY=1
run PZ_GRID_NO_OVERLAP=1
run NCCL_MAX_CTAS=4
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=16
run NCCL_MAX_CTAS=4 PZ_GRID_NO_OVERLAP=1
cat gpurun_out/r2x_sweep.txt
true
