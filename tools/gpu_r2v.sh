#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
tail -3 gpurun_out/r2v_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2v_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d.get('cpu_reference_forward'), d['roofline']['frac'], d['roofline'].get('peak_note'))"
true
