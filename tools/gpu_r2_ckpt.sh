#!/bin/bash
# checkpoint: full GPU suite + default bench (N=1) + reference arm
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/ck_pytest.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/ck_pytest.log
timeout 1500 python bench.py > gpurun_out/ck_bench_n1.json 2> gpurun_out/ck_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/ck_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/ck_bench_n1.json"))
print("value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d["single_proof_latency_ms"]))
print({k:(round(v["ms_per_launch"],3), v["launches"]) for k,v in d["stages"].items()})
w=d["witness"]; print("witness gen+check %.0f  gen %.0f  sat %.0f /s"%(w["value"],w["generate_only"],w["satisfy_only"]))
print({k:(v.get("value"), v.get("seconds")) for k,v in d["extra"].items()}); print(d["cpu_baseline"]["value"])
PY
