#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_abi.py -x -q -m gpu > gpurun_out/sqr_abi.log 2>&1; echo "abi rc=$?"; tail -15 gpurun_out/sqr_abi.log
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_split.py tests/test_gpu_ntt.py tests/test_gpu_setup.py -x -q -m gpu > gpurun_out/sqr_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/sqr_pytest.log
bash tools/gpu_r2_pf.sh sqr=FRCS_MSM_PF=0 sqr_pair=FRCS_MSM_PAIR=1 sqr_pair3=FRCS_MSM_PAIR=1,FRCS_MSM_PAIR_LEVELS=3 nowave=FRCS_MSM_ONE_WAVE=0
