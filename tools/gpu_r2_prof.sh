#!/bin/bash
# ncu launch list of a small bench run (one group of 16 proofs per step), after the same command ran clean
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 64 --no-cpu-baseline"
timeout 600 $SMALL > gpurun_out/p_plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches.csv $SMALL > gpurun_out/p_ncu_list.log 2>&1
echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/p_launches.csv --group > gpurun_out/p_summary.txt 2>&1
head -60 gpurun_out/p_summary.txt
