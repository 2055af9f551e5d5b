#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_nets.py -m gpu -q --tb=short -k "16bit" 2>&1 | tail -40
true
