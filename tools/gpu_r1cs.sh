#!/bin/bash
# R1CS iteration: GPU parity tests, bench, launch list of the witness + satisfaction path, full capture of the R1CS kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_it.json 2> gpurun_out/bench_it.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_it.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_it.json"))
    print("value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"]))
    print({k:round(v["ms_per_launch"],3) for k,v in d["stages"].items()})
    print("witness", {k:v for k,v in d["witness"].items() if k!="roofline"}, d["witness"]["roofline"]["frac"])
except Exception as e: print("no json", e)
PY
timeout 600 python tools/prof_witness.py 592 > gpurun_out/plain_w.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_w.csv python tools/prof_witness.py 592 > gpurun_out/ncu_w.log 2>&1
python tools/launch_summary.py gpurun_out/launches_w.csv 2>/dev/null | head -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:witness_kernel\|r1cs_ -s 4 -c 4 -o gpurun_out/prof_witness python tools/prof_witness.py 592 > gpurun_out/ncu_full_w.log 2>&1
echo "ncu witness rc=$?"
