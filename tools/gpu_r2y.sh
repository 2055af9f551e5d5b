#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seam_reference_unittests.py -m gpu -q --tb=short -k "PRelu or Pad1D or Pad2D or Upsample or LCN or MapLRN or side_modules or CuDnnNorm" 2>&1 | tail -60 > gpurun_out/r2y_pytest.log
tail -n 60 gpurun_out/r2y_pytest.log
true
