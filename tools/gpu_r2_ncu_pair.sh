#!/bin/bash
# ncu --set full of the batched-affine pair levels (TABLE and array variants), one group of 16 proofs
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
timeout 600 $SMALL > gpurun_out/n_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/n_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pair_kernel|accumA_kernel" -s 9 -c 3 -o gpurun_out/n_pair $SMALL > gpurun_out/n_pair.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/n_pair.ncu-rep --page details > gpurun_out/n_pair_details.txt 2>&1
grep -E "^  [a-zA-Z<]|Duration|DRAM Throughput|Compute \(SM\) Throughput|Registers Per|Achieved Occupancy|Issued Ipc Active|Avg. Active Threads|L2 Hit|L1/TEX Hit|Stall|stall" gpurun_out/n_pair_details.txt | cut -c1-220 | head -60
ncu -i gpurun_out/n_pair.ncu-rep --page raw --csv --metrics smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum > gpurun_out/n_pair_stalls.csv 2>&1
cat gpurun_out/n_pair_stalls.csv | cut -c1-600
