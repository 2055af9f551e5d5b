#!/bin/bash
# MN-major activation operand through the copy engine (MODE_MN_TMA): descriptor variants (PZ_DEBUG_SKIP bits 8 = LBO/SBO swapped, 16 = SBO 1024)
mkdir -p gpurun_out
{
for v in 0 8 16 24; do
PZ_TMA_FPROP=2 PZ_DEBUG_SKIP=$v timeout 120 python tools/check_tma_fprop.py 2>&1 | tail -10
done
} > gpurun_out/r4c_check.txt 2>&1
cat gpurun_out/r4c_check.txt
true
