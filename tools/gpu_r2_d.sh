#!/bin/bash
# round 2, call D: gadget entry points on CUDA (reference KATs), sort-span diagnostics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gadgets.py tests/test_abi.py -x -q --durations=5 > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/d_pytest.log
run_bench() {
  name=$1; shift
  env $FRCS_ENV timeout 900 python bench.py --no-cpu-baseline "$@" > gpurun_out/d_bench_$name.json 2> gpurun_out/d_bench_$name.err; echo "bench $name rc=$?"
  tail -2 gpurun_out/d_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/d_bench_$name.json"))
    print("$name value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d["single_proof_latency_ms"]))
    print({k:(round(v["ms_per_launch"],3), v["launches"]) for k,v in d["stages"].items()})
except Exception as e: print("no json", e)
PY
}
run_bench b64 --steps 4 --warmup 3 --batch 64
