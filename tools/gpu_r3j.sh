#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 3 -c 1 -o gpurun_out/r02c_l4_fprop python tools/bench_layers.py 64 4 > gpurun_out/r02c_l4.log 2>&1
ncu -i gpurun_out/r02c_l4_fprop.ncu-rep --page source --csv > gpurun_out/r02c_l4_fprop_source.csv 2>/dev/null
ncu -i gpurun_out/r02c_l4_fprop.ncu-rep --page raw --csv > gpurun_out/r02c_l4_fprop_raw.csv 2>/dev/null
rm -f gpurun_out/r02c_l4_fprop.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 15 -c 1 -o gpurun_out/r02c_l4_wgrad python tools/bench_layers.py 64 4 > gpurun_out/r02c_l4w.log 2>&1
ncu -i gpurun_out/r02c_l4_wgrad.ncu-rep --page source --csv > gpurun_out/r02c_l4_wgrad_source.csv 2>/dev/null
ncu -i gpurun_out/r02c_l4_wgrad.ncu-rep --page raw --csv > gpurun_out/r02c_l4_wgrad_raw.csv 2>/dev/null
rm -f gpurun_out/r02c_l4_wgrad.ncu-rep
ls -la gpurun_out/r02c_*
true
