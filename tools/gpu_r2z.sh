#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gen_golden_cuda.py --impl ref --side --out gpurun_out/ref_cuda_side.npz > gpurun_out/r2z_gen_ref.log 2>&1
timeout 600 python tools/gen_golden_cuda.py --impl b200 --side --out gpurun_out/b200_side.npz > gpurun_out/r2z_gen_b200.log 2>&1
tail -3 gpurun_out/r2z_gen_ref.log; tail -3 gpurun_out/r2z_gen_b200.log
python - <<'PY'
import numpy as np
a = np.load("gpurun_out/ref_cuda_side.npz"); b = np.load("gpurun_out/b200_side.npz")
for k in a.files:
    if k not in b.files: print("missing", k); continue
    x, y = a[k].astype(np.float64), b[k].astype(np.float64)
    if x.shape != y.shape: print("shape", k, x.shape, y.shape); continue
    err = np.abs(x - y).max() / (np.abs(x).max() + 1e-30)
    print("%-40s %.3e %s" % (k, err, "" if err < 2e-5 else "<<<<"))
PY
true
