#!/bin/bash
mkdir -p gpurun_out
run_bench() {
  name=$1; shift
  env $FRCS_ENV timeout 900 python bench.py --no-cpu-baseline "$@" > gpurun_out/e_bench_$name.json 2> gpurun_out/e_bench_$name.err; echo "bench $name rc=$?"
  tail -2 gpurun_out/e_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e_bench_$name.json"))
    print("$name value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d["single_proof_latency_ms"]))
    print({k:(round(v["ms_per_launch"],3), v["launches"]) for k,v in d["stages"].items()})
except Exception as e: print("no json", e)
PY
}
run_bench b32 --steps 4 --warmup 3 --batch 32


