#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_split.py tests/test_gpu_schoolbook.py tests/test_gpu_dual.py -x -q -m gpu > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/fin_pytest.log
bash tools/gpu_r2_pf.sh main=FRCS_MSM_PF=0
