#!/bin/bash
# round 2, call A: GPU parity tests (incl. the new F1024 large-batch / 1000-signature / SB1024 split tests and both MSM
# window geometries), then bench A/B: narrow (8-bit) vs wide (16-bit) geometry for the z MSMs, and the G2 variant.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/a_pytest.log
run_bench() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_$name.json 2> gpurun_out/a_bench_$name.err; echo "bench $name rc=$?"
  tail -2 gpurun_out/a_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/a_bench_$name.json"))
    print("$name value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d["single_proof_latency_ms"]))
    print({k:round(v["ms_per_launch"],3) for k,v in d["stages"].items()})
    print("witness", d["witness"]["value"], d["witness"]["generate_only"], d["witness"]["satisfy_only"])
except Exception as e: print("no json", e)
PY
}
run_bench narrow FRCS_X=0
run_bench wide FRCS_Z_WINDOW_BITS=16
run_bench g2b2 FRCS_LIB=$PWD/falcon_r1cs_b200/variants/g2b2.so
