#!/bin/bash
# split-proof iteration: GPU tests of the split entry points (N=1 part), then the latency variants under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py -x -q -m gpu > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s_pytest.log
shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/split_latency.py "$@" > gpurun_out/s_split_n$N.jsonl 2> gpurun_out/s_split_n$N.err; echo "split rc=$?"; tail -5 gpurun_out/s_split_n$N.err | cut -c1-400
cat gpurun_out/s_split_n$N.jsonl
