#!/bin/bash
# multi-GPU evidence (N = $1): bench, config 3, split schoolbook proof
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err
timeout 600 $TR tools/run_config3.py > gpurun_out/config3_n$N.json 2>/dev/null; tail -1 gpurun_out/config3_n$N.json
timeout 900 $TR tools/bench_split.py --kind 1 --logn 10 --steps 3 --warmup 1 > gpurun_out/split_sb_n$N.json 2>/dev/null; tail -1 gpurun_out/split_sb_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print("N=$N value %.1f e2e %.1f witness %.0f gen %.0f sat %.0f"%(d["value"],d["e2e"]["value"],d["witness"]["value"],d["witness"]["generate_only"],d["witness"]["satisfy_only"]))
PY
