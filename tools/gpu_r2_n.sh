#!/bin/bash
# launch list of the current default path (one proof group of 16) + a short bench line
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
timeout 600 $SMALL > gpurun_out/n_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/n_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_launches.csv $SMALL > gpurun_out/n_ncu_list.log 2>&1
echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/n_launches.csv > gpurun_out/n_launches_summary.txt; head -60 gpurun_out/n_launches_summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/n_bench.json"))
print("value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d["single_proof_latency_ms"]))
print({k:(round(v["ms_per_launch"],3), v["launches"]) for k,v in d["stages"].items()})
PY
