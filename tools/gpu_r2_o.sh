#!/bin/bash
# group-size sweep: proofs/s with FRCS_GROUP = 16 / 32 / 64 at 64 and 128 proofs per step
mkdir -p gpurun_out
show() { python - "$1" <<PY
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.1f e2e %.1f ms/step %.1f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["single_proof_latency_ms"]))
print({k:round(v["ms_per_launch"],2) for k,v in d["stages"].items()})
PY
}
for cfg in "32 64" "64 64" "32 128" "16 128"; do
  set -- $cfg
  FRCS_GROUP=$1 timeout 600 python bench.py --steps 3 --warmup 3 --batch $2 --no-cpu-baseline --no-extra > gpurun_out/o_g$1_b$2.json 2> gpurun_out/o_g$1_b$2.err || { echo "failed $cfg"; tail -5 gpurun_out/o_g$1_b$2.err; continue; }
  show gpurun_out/o_g$1_b$2.json
done
