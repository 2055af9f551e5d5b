#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/run_ref_unittests.py --impl b200 --out gpurun_out/r3a_ref_unittests_b200.json > gpurun_out/r3a_ref_unittests.log 2>&1
grep -v " ok$" gpurun_out/r3a_ref_unittests.log | tail -40
true
