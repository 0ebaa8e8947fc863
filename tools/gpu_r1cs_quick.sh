#!/bin/bash
# quick R1CS iteration: parity tests of the witness / R1CS path, warm timing, per-kernel launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_witness.py tests/test_gpu_schoolbook.py -m gpu -x -q 2>&1 | tail -2
python tools/time_r1cs.py 592
python tools/time_r1cs.py 592 9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_w.csv python tools/prof_witness.py 592 > gpurun_out/ncu_w.log 2>&1
python tools/launch_summary.py gpurun_out/launches_w.csv 2>/dev/null | head -8
