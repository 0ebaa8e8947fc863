#!/bin/bash
# witness generation overlapped with the satisfaction check (two streams): parity, then witnesses/s and config 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_witness.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/x_new.json 2> gpurun_out/x_new.err
python - <<PY
import json
d=json.load(open("gpurun_out/x_new.json"))
w=d["witness"]
print("value %.1f | wit %.0f sat %.0f gen %.0f | config3 %.0f /s %.4f s"%(d["value"],w["value"],w["satisfy_only"],w["generate_only"],d["extra"]["config3"]["value"],d["extra"]["config3"]["seconds"]))
PY
FRCS_CHECK_OVERLAP=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/x_old.json 2> gpurun_out/x_old.err
python - <<PY
import json
d=json.load(open("gpurun_out/x_old.json"))
w=d["witness"]
print("no overlap: wit %.0f | config3 %.0f /s"%(w["value"],d["extra"]["config3"]["value"]))
PY
