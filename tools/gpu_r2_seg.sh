#!/bin/bash
# segmented level-0 pieces: MSM / proof parity, then the bench with slice-length variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_split.py tests/test_gpu_setup.py -x -q -m gpu > gpurun_out/seg_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/seg_pytest.log
for v in main lc64 lc16; do
  if [ $v != main ]; then export FRCS_LIB=$PWD/falcon_r1cs_b200/variants/$v.so; else unset FRCS_LIB; fi
  timeout 600 python bench.py --no-cpu-baseline --no-extra --steps 4 --warmup 3 > gpurun_out/seg_bench_$v.json 2> gpurun_out/seg_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/seg_bench_$v.json"))
    print("$v value %.1f e2e %.1f  ms/step %.1f roof %.3f accum %.2f msm_h %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["frac"],d["stages"]["msm_h_accum"]["ms_per_launch"],d["stages"]["msm_h"]["ms_per_launch"]))
except Exception as e: print("no json", e)
PY
done
