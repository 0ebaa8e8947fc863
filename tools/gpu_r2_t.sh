#!/bin/bash
# satisfaction check timing (device-resident z) + witness parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_witness.py tests/test_gpu_dual.py tests/test_gpu_schoolbook.py tests/test_gpu_gadgets.py -m gpu -x -q 2>&1 | tail -4
for n in 592 4096; do timeout 300 python tools/time_r1cs.py $n 10 20 2>&1 | tail -1; done
timeout 300 python tools/time_r1cs.py 592 9 20 2>&1 | tail -1
