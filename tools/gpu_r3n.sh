#!/bin/bash
for l in 13 14 18 19 1; do python tools/bench_layers.py 1 $l 2>&1 | grep -v "^layer\|^sum"; done
for l in 13 19; do python tools/bench_layers.py 8 $l 2>&1 | grep -v "^layer\|^sum"; done
true
