#!/bin/bash
N=${1:-8}; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/split_latency.py "$@" > gpurun_out/s_split_n$N.jsonl 2> gpurun_out/s_split_n$N.err; echo "split rc=$?"; tail -3 gpurun_out/s_split_n$N.err | cut -c1-300
cat gpurun_out/s_split_n$N.jsonl
