#!/bin/bash
# full default bench at N=1 (with the config3 / Falcon-512 blocks and the CPU baseline) + reference arm
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/i_bench_n1.json 2> gpurun_out/i_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/i_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/i_bench_n1.json"))
print("value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d["single_proof_latency_ms"]))
print({k:(round(v["ms_per_launch"],3), v["launches"]) for k,v in d["stages"].items()})
w=d["witness"]; print("witness gen+check %.0f  gen %.0f  sat %.0f /s"%(w["value"],w["generate_only"],w["satisfy_only"]))
print(json.dumps(d["extra"],indent=1)); print(d["cpu_baseline"]); print(d["roofline"].get("textbook_normalised"))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/i_bench_ref.json 2> gpurun_out/i_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/i_bench_ref.json | cut -c1-600
