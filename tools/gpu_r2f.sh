#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x -k "batchnorm or deferred or parity or seam" 2>&1 | tail -60 > gpurun_out/r2f_pytest.log
for kb in 48 96 192; do for l2 in 48 80 110; do for th in 256 512; do
  echo "== CTA_KB=$kb L2_MB=$l2 THREADS=$th" >> gpurun_out/r2f_bn_sweep.log
  PZ_BN_CTA_KB=$kb PZ_BN_L2_MB=$l2 PZ_BN_THREADS=$th timeout 120 python tools/bench_ops.py 64 bn 2>&1 | grep -E "bn_|totals" >> gpurun_out/r2f_bn_sweep.log
done; done; done
tail -n 4 gpurun_out/r2f_pytest.log; grep -E "==|totals" gpurun_out/r2f_bn_sweep.log
true
