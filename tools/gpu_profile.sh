#!/bin/bash
# ncu evidence of one round (B200_PROFILING.md): the launch list of a bench step with DRAM counters, and --set full captures of
# the dominant kernels.  usage (under gpurun): bash tools/gpu_profile.sh r02
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-run --warmup 1 --steps 1 > gpurun_out/${tag}_launches.log 2>&1
# layer 3 of tools/bench_layers.py: 64 -> 256 channels 1x1 at 55x55, N = 64 (the write-heavy HBM-bound conv): fprop launch
ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 3 -c 1 -o gpurun_out/${tag}_l3_fprop \
    python tools/bench_layers.py 64 3 > gpurun_out/${tag}_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_fwd_cluster -s 1 -c 1 -o gpurun_out/${tag}_bn_fwd \
    python tools/bench_ops.py 64 256x55 > gpurun_out/${tag}_bn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_bwd_cluster -s 1 -c 1 -o gpurun_out/${tag}_bn_bwd \
    python tools/bench_ops.py 64 512x28 >> gpurun_out/${tag}_bn.log 2>&1
ls -la gpurun_out/${tag}_*
