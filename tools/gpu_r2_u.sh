#!/bin/bash
# six-transform witness map: parity (witness map, proofs of every circuit, split key), then proofs/s
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_prove.py tests/test_gpu_dual.py tests/test_gpu_schoolbook.py tests/test_gpu_split.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/u_new.json 2> gpurun_out/u_new.err
python - <<PY
import json
d=json.load(open("gpurun_out/u_new.json"))
print("value %.1f e2e %.1f ms/step %.1f lat %.2f roof %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["single_proof_latency_ms"],d["roofline"]["frac"]))
print({k:round(v["ms_per_launch"],2) for k,v in d["stages"].items()})
print([ (r["kernel"][:20], round(r["frac"],3)) for r in d["rooflines"]])
PY
