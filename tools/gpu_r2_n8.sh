#!/bin/bash
# 8-GPU evidence: bench.py under torchrun (weak scaling, config 3, split schoolbook proof in `extra`)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 8 --no-cpu-baseline > gpurun_out/f_bench_n8.json 2> gpurun_out/f_bench_n8.err; echo "bench rc=$?"; tail -2 gpurun_out/f_bench_n8.err
python - <<PY
import json
d=json.load(open("gpurun_out/f_bench_n8.json"))
print("N=8 value %.1f e2e %.1f witness %.0f"%(d["value"],d["e2e"]["value"],d["witness"]["value"]))
print(json.dumps(d["extra"])[:2500])
PY
