#!/bin/bash
mkdir -p gpurun_out
FRCS_DEBUG_BENCH=1 timeout 900 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; grep config3 gpurun_out/j_bench.err
python -c "
import json; d=json.load(open('gpurun_out/j_bench.json')); print(d['extra']['config3']['seconds'], d['value'])"
