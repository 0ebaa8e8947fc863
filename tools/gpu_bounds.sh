#!/bin/bash
# The GPU parity suite against a build with device-side bounds asserts (-DFRCS_BOUNDS): our substitute for
# compute-sanitizer memcheck, which is closed on this GPU pool (profiles/r02_compute_sanitizer_attempt.txt).
mkdir -p gpurun_out
export FRCS_LIB=$PWD/falcon_r1cs_b200/variants/bounds.so
timeout 1800 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/bounds_pytest.log 2>&1; echo "pytest (bounds build) rc=$?"; tail -4 gpurun_out/bounds_pytest.log
python tools/sanitize_check.py 9 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
