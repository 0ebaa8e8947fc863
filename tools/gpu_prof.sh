#!/bin/bash
# ncu launch list of a small bench run (group of 16), after the same command ran clean
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --batch 16 --wbatch 16 --no-cpu-baseline"
timeout 600 $SMALL > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
