#!/bin/bash
# round 2, call F: streaming short-row kernel (TMA-staged windows): parity, then witnesses/s
mkdir -p gpurun_out
FRCS_DEBUG=1 timeout 1200 python -m pytest tests/test_gpu_witness.py tests/test_gpu_dual.py -x -q -m "gpu" --durations=5 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/f_pytest.log; grep "stream plan" gpurun_out/f_pytest.log | sort | uniq -c
for v in stream nostream; do
  if [ $v = nostream ]; then export FRCS_NO_STREAM=1; fi
  timeout 900 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --batch 32 > gpurun_out/f_bench_$v.json 2> gpurun_out/f_bench_$v.err; echo "bench $v rc=$?"
  tail -2 gpurun_out/f_bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/f_bench_$v.json"))
    print("$v value %.1f e2e %.1f proofs/s  ms/step %.1f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]))
    print({k:(round(x["ms_per_launch"],3), x["launches"]) for k,x in d["stages"].items() if k in ("r1cs","witness","witness_map")})
    w=d["witness"]; print("witness gen+check %.0f  gen %.0f  sat %.0f /s  sat frac %.3f"%(w["value"],w["generate_only"],w["satisfy_only"],w["satisfy_roofline"]["frac"]))
except Exception as e: print("no json", e)
PY
done
