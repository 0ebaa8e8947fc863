#!/bin/bash
# quick iteration: GPU tests + bench at several group sizes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for g in ${GROUPS_TO_RUN:-16}; do
  FRCS_GROUP=$g timeout 900 python bench.py --batch ${BATCH:-32} --steps ${STEPS:-4} --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err; echo "bench g=$g rc=$?"
  tail -3 gpurun_out/bench_g$g.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_g$g.json"))
    print("g=$g value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"]))
    print({k:round(v["ms_per_launch"],3) for k,v in d["stages"].items()})
    print("witness", d["witness"]["value"], d["witness"]["generate_only"], d["witness"]["roofline"]["frac"])
except Exception as e: print("no json", e)
PY
done
