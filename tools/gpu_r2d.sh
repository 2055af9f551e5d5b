#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_seam_reference_unittests.py --tb=short 2>&1 | tail -150 > gpurun_out/r2d_pytest.log
timeout 600 python -m pytest tests/test_gpu_seam_reference_unittests.py -q --tb=short -k "Activation or BatchNorm1D or BatchNorm3D" 2>&1 | tail -120 > gpurun_out/r2d_pytest_seam.log
tail -n 3 gpurun_out/r2d_pytest.log gpurun_out/r2d_pytest_seam.log
true
