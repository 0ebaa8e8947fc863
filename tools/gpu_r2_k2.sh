#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_split.py tests/test_gpu_schoolbook.py tests/test_gpu_dual.py -x -q -m gpu > gpurun_out/k2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k2_pytest.log
bash tools/gpu_r2_k.sh $N 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep -A40 "^N="
