#!/bin/bash
# schoolbook parity after the multi-CTA witness split, then compute-sanitizer memcheck / racecheck on smoke() and on a
# witness_check_batch of 128
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schoolbook.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_sanitizer_${tool}_smoke.log 2>&1; echo "$tool smoke rc=$?"
  tail -4 gpurun_out/l_sanitizer_${tool}_smoke.log
  timeout 1500 $CS --tool $tool --print-limit 20 python tools/sanitize_check.py 9 > gpurun_out/l_sanitizer_${tool}_check.log 2>&1; echo "$tool check rc=$?"
  tail -4 gpurun_out/l_sanitizer_${tool}_check.log
done
