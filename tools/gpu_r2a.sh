#!/bin/bash
# round 2, call A: reference CUDA backend as oracle + perf anchor; the reference's own unit tests over the seam
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r2a_env.log 2>&1
timeout 600 python tools/gen_golden_cuda.py --impl ref --out gpurun_out/ref_cuda_ops.npz --nets gpurun_out/ref_cuda_nets.npz > gpurun_out/r2a_golden_ref.log 2>&1
timeout 600 python tools/gen_golden_cuda.py --impl b200 --out gpurun_out/b200_ops.npz --nets gpurun_out/b200_nets.npz > gpurun_out/r2a_golden_b200.log 2>&1
timeout 600 python tools/bench_ref_cuda.py --impl ref --model resnet50 --batch 64 > gpurun_out/r2a_bench_ref_r50.log 2>&1
timeout 600 python tools/bench_ref_cuda.py --impl b200 --model resnet50 --batch 64 > gpurun_out/r2a_bench_b200_r50.log 2>&1
timeout 600 python tools/bench_ref_cuda.py --impl ref --model resnet50 --batch 64 --forward-only > gpurun_out/r2a_bench_ref_r50_fwd.log 2>&1
timeout 600 python tools/bench_ref_cuda.py --impl ref --model vgg16 --batch 128 --dtype f16 --steps 5 --warmup 3 > gpurun_out/r2a_bench_ref_vgg_f16.log 2>&1
timeout 600 python tools/bench_ref_cuda.py --impl ref --model lenet --batch 64 > gpurun_out/r2a_bench_ref_lenet.log 2>&1
timeout 1200 python tools/run_ref_unittests.py --impl b200 --out gpurun_out/r2a_unittests_b200.json > gpurun_out/r2a_unittests_b200.log 2>&1
timeout 1200 python tools/run_ref_unittests.py --impl ref --out gpurun_out/r2a_unittests_ref.json > gpurun_out/r2a_unittests_ref.log 2>&1
for f in gpurun_out/r2a_bench*.log; do tail -n 2 $f; done; true
timeout 600 python tools/bench_ref_cuda.py --impl b200 --model vgg16 --batch 128 --dtype f16 --steps 5 --warmup 3 > gpurun_out/r2a_bench_b200_vgg_f16.log 2>&1
timeout 600 python tools/bench_ref_cuda.py --impl b200 --model lenet --batch 64 > gpurun_out/r2a_bench_b200_lenet.log 2>&1
