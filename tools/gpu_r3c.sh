#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "lstm or rnn or gru" 2>&1 | tail -15
true
