#!/bin/bash
# ncu: launch list of witness + satisfaction at 592 signatures, full capture of the streaming short-row kernel
mkdir -p gpurun_out
timeout 600 python tools/prof_witness.py 592 > gpurun_out/g_plain_w.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches_w.csv python tools/prof_witness.py 592 > gpurun_out/g_ncu_w.log 2>&1
python tools/launch_summary.py gpurun_out/g_launches_w.csv 2>/dev/null | head -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:r1cs_stream -s 1 -c 1 -o gpurun_out/g_prof_stream python tools/prof_witness.py 592 > gpurun_out/g_ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/g_prof_stream.ncu-rep --page details > gpurun_out/g_stream_details.txt 2>&1
grep -E "Duration|DRAM Throughput|Memory Throughput|L1/TEX Hit|Registers Per|Achieved Occupancy|Theoretical Occupancy|Issued Ipc|Eligible|No Eligible|Shared Memory Configuration|Dynamic Shared|Block Limit|Bank" gpurun_out/g_stream_details.txt | head -40
