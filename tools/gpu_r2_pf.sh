#!/bin/bash
# software-prefetch variants of the accumulation kernels (per-bucket pieces and batched-affine pair levels)
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-extra --steps 4 --warmup 3 > gpurun_out/pf_bench_$name.json 2> gpurun_out/pf_bench_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pf_bench_$name.json"))
    print("$name value %.1f e2e %.1f  ms/step %.1f accum %.2f msm_h %.2f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["stages"]["msm_h_accum"]["ms_per_launch"],d["stages"]["msm_h"]["ms_per_launch"],d["single_proof_latency_ms"]))
except Exception as e: print("no json", e)
PY
}
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}
  run $name ${envs//,/ }
done
