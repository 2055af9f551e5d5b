#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --tb=short -k "checkpoint" 2>&1 | tail -15 > gpurun_out/r2t_pytest.log
for cfg in "PZ_BN_STASH_KB=100 PZ_BN_MAX_CL=8" "PZ_BN_STASH_KB=200 PZ_BN_MAX_CL=8" "PZ_BN_STASH_KB=100 PZ_BN_MAX_CL=16" "PZ_BN_STASH_KB=200 PZ_BN_MAX_CL=16"; do
  echo "== $cfg" >> gpurun_out/r2t_bn.txt
  env $cfg timeout 600 python tools/bench_ops.py 64 bn 2>&1 | grep "112x112\|55x55\|totals" >> gpurun_out/r2t_bn.txt
  env $cfg timeout 900 python bench.py --steps 10 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']
print(d['ms_per_step'], 'bn_fwd', f['bn_fwd'], 'bn_bwd', f['bn_bwd'])" >> gpurun_out/r2t_bn.txt
done
tail -n 5 gpurun_out/r2t_pytest.log; cat gpurun_out/r2t_bn.txt
true
