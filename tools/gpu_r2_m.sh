#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_witness.py tests/test_gpu_dual.py tests/test_gpu_prove.py tests/test_gpu_gadgets.py -x -q -m "gpu and not slow" --durations=4 > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/m_pytest.log
timeout 300 python tools/time_sat.py 592 2>&1 | tail -1
timeout 300 python tools/time_sat.py 4096 2>&1 | tail -1
FRCS_NO_NTT_ROWS=1 timeout 300 python tools/time_sat.py 592 2>&1 | tail -1
timeout 300 python tools/time_sat.py 592 9 2>&1 | tail -1
timeout 600 python tools/run_config3.py 2>/dev/null
