#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_witness.py tests/test_gpu_dual.py -x -q -m "gpu and not slow" > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/h_pytest.log
for v in main sw1024c3 sw512c4; do
  if [ $v != main ]; then export FRCS_LIB=$PWD/falcon_r1cs_b200/variants/$v.so; fi
  timeout 300 python tools/time_sat.py 592 2>&1 | tail -1
  timeout 300 python tools/time_sat.py 4096 2>&1 | tail -1
done
unset FRCS_LIB
FRCS_NO_STREAM=1 timeout 300 python tools/time_sat.py 592 2>&1 | tail -1
