#!/bin/bash
# ncu evidence of the FINAL code of a round, exported as text on the box (the .ncu-rep files exceed what gpurun brings back)
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-run --warmup 1 --steps 1 > gpurun_out/${tag}_launches.log 2>&1
cap() {   # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/${tag}_$name "$@" > gpurun_out/${tag}_$name.log 2>&1
  ncu -i gpurun_out/${tag}_$name.ncu-rep --page details > gpurun_out/${tag}_${name}_ncu_details.txt 2>/dev/null
  python tools/ncu_keys.py gpurun_out/${tag}_$name.ncu-rep > gpurun_out/${tag}_${name}_ncu_keys.txt 2>/dev/null
  rm -f gpurun_out/${tag}_$name.ncu-rep
}
cap l3_fprop umma_gemm 3 python tools/bench_layers.py 64 3
cap bn_fwd bn_fwd_cluster 1 python tools/bench_ops.py 64 256x55
cap bn_bwd bn_bwd_cluster 1 python tools/bench_ops.py 64 512x28
ls -la gpurun_out/${tag}_*
