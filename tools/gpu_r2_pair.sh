#!/bin/bash
# batched-affine pair levels: MSM parity with the pair levels forced at every size, proof parity, then bench variants
mkdir -p gpurun_out
FRCS_MSM_PAIR=2 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu > gpurun_out/pair_pytest_forced.log 2>&1; echo "pytest forced rc=$?"; tail -4 gpurun_out/pair_pytest_forced.log
timeout 1200 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/pair_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pair_pytest.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-extra --steps 4 --warmup 3 > gpurun_out/pair_bench_$name.json 2> gpurun_out/pair_bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pair_bench_$name.json"))
    print("$name value %.1f e2e %.1f  ms/step %.1f accum %.2f msm_h %.2f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["stages"]["msm_h_accum"]["ms_per_launch"],d["stages"]["msm_h"]["ms_per_launch"],d["single_proof_latency_ms"]))
except Exception as e: print("no json", e)
PY
}
run off FRCS_MSM_PAIR=0
run p2m256 FRCS_MSM_PAIR=1
run p3m256 FRCS_MSM_PAIR_LEVELS=3
run p2m128 FRCS_MSM_PAIR_M=128
run p2m512 FRCS_MSM_PAIR_M=512
run p1m256 FRCS_MSM_PAIR_LEVELS=1
