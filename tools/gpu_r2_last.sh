#!/bin/bash
# last check of the round: GPU suite and smoke on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/f_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
