#!/bin/bash
# ncu launch list of the final code (2 eager steps); one full capture of a copied 1x1 wgrad if time allows
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02g_launches.csv python bench.py --profile-run --warmup 1 --steps 1 > gpurun_out/r02g_launches.log 2>&1
ls -la gpurun_out/r02g_launches.csv
timeout 45 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 2 -c 1 -o gpurun_out/r02g_l8_wgrad python tools/bench_layers.py 64 8 > gpurun_out/r02g_l8_wgrad.log 2>&1
ncu -i gpurun_out/r02g_l8_wgrad.ncu-rep --page details > gpurun_out/r02g_l8_wgrad_ncu_details.txt 2>/dev/null
python tools/ncu_keys.py gpurun_out/r02g_l8_wgrad.ncu-rep > gpurun_out/r02g_l8_wgrad_ncu_keys.txt 2>/dev/null
rm -f gpurun_out/r02g_l8_wgrad.ncu-rep
true
