#!/bin/bash
mkdir -p gpurun_out
PZ_TMA_WGRAD=3 timeout 300 compute-sanitizer --print-limit 5 python tools/check_tma_wgrad.py f32 4 > gpurun_out/r4g_sanitizer.txt 2>&1
head -60 gpurun_out/r4g_sanitizer.txt
true
