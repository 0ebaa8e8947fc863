#!/bin/bash
# N-GPU bench (driver's launch line) incl. the split schoolbook proof
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/k_bench_n$N.json 2> gpurun_out/k_bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/k_bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/k_bench_n$N.json"))
print("N=%d value %.1f e2e %.1f proofs/s  ms/step %.1f"%(d["n_gpus"],d["value"],d["e2e"]["value"],d["ms_per_step"]))
w=d["witness"]; print("witness gen+check %.0f  gen %.0f  sat %.0f /s"%(w["value"],w["generate_only"],w["satisfy_only"]))
print(json.dumps(d["extra"],indent=1))
PY
