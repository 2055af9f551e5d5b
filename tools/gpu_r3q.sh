#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --tb=short -x -k "batchnorm or bn or parity or nets or instancenorm or BatchNorm or CuDnnNorm or fullsize or folded" 2>&1 | tail -6
true
