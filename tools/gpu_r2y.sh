#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seam_reference_unittests.py tests/test_gpu_parity_cuda.py -m gpu -q --tb=short -k "CTC or side_module" 2>&1 | tail -60 > gpurun_out/r2y_pytest.log
tail -n 60 gpurun_out/r2y_pytest.log
true
