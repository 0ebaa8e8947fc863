#!/bin/bash
# 2-GPU bench line on the final tree
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/f_bench_n2.json 2> gpurun_out/f_bench_n2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/f_bench_n2.json"))
print("N=2 value %.1f e2e %.1f witness %.0f split %.2f ms config3 %.4f s"%(d["value"],d["e2e"]["value"],d["witness"]["value"],d["extra"]["split"]["ms_per_proof"],d["extra"]["config3"]["seconds"]))
PY
