#!/bin/bash
mkdir -p gpurun_out
for l in 3 4 8 13 14 2 7; do
  PZB200_LIB=$PWD/puzzlelib_b200/libpzb200_timeline.so timeout 300 python tools/bench_layers.py 64 $l >> gpurun_out/r2q_timeline.txt 2>&1
done
timeout 600 python tools/bench_layers.py > gpurun_out/r2q_layers.txt 2>&1
timeout 600 python tools/bench_ops.py 64 bn > gpurun_out/r2q_bn.txt 2>&1
tail -n 2 gpurun_out/r2q_layers.txt; tail -n 30 gpurun_out/r2q_bn.txt
true
