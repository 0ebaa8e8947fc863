#!/bin/bash
# ncu launch list of witness generation + satisfaction at 592 signatures
mkdir -p gpurun_out
timeout 600 python tools/prof_witness.py 592 > gpurun_out/g_plain_w.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches_w.csv python tools/prof_witness.py 592 > gpurun_out/g_ncu_w.log 2>&1
python tools/launch_summary.py gpurun_out/g_launches_w.csv 2>/dev/null | head -12
timeout 600 python tools/run_config3.py 2>/dev/null
timeout 600 python tools/run_config3.py 2>/dev/null
