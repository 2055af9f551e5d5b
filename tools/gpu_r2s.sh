#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/measure_tf32_peak.py gpurun_out/r2s_tf32_peak.json > gpurun_out/r2s_tf32.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2s_pytest.log
cat gpurun_out/r2s_tf32.log | tail -2; tail -n 40 gpurun_out/r2s_pytest.log
true
