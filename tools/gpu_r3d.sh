#!/bin/bash
# side configurations (BASELINE.json configs 3..5) and the final ncu evidence of round 2
mkdir -p gpurun_out
timeout 900 python bench.py --model vgg16 --dtype f16 --batch 128 --steps 10 --warmup 3 --no-cpu > gpurun_out/r3d_bench_vgg16_f16.json 2> gpurun_out/r3d_vgg_f16.err
timeout 900 python bench.py --model vgg16 --dtype bf16 --batch 128 --steps 10 --warmup 3 --no-cpu --no-ref-gpu > gpurun_out/r3d_bench_vgg16_bf16.json 2> gpurun_out/r3d_vgg_bf16.err
timeout 900 python bench.py --model resnet50 --dtype bf16 --steps 20 --warmup 5 --no-cpu --no-ref-gpu > gpurun_out/r3d_bench_resnet50_bf16.json 2> gpurun_out/r3d_r50_bf16.err
timeout 600 python tools/bench_lstm.py > gpurun_out/r3d_lstm_config5.json 2> gpurun_out/r3d_lstm.err
for f in vgg16_f16 vgg16_bf16 resnet50_bf16; do python -c "
import json
d=json.loads(open('gpurun_out/r3d_bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), round(d['ms_per_step'],2), d.get('reference_gpu',{}).get('value'), d['roofline']['kernel'][:30], round(d['roofline']['frac'],3))"; done
tail -2 gpurun_out/r3d_lstm_config5.json | head -c 600
bash tools/gpu_profile.sh r02b > gpurun_out/r3d_profile.log 2>&1
tail -5 gpurun_out/r3d_profile.log
true
