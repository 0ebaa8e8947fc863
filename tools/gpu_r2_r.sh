#!/bin/bash
# block-histogram sort of the l+h scalars: parity, proofs/s against the per-entry atomics, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_dual.py -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r_pytest.log
show() { python - "$1" <<PY
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.1f e2e %.1f ms/step %.1f lat %.2f roof %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["single_proof_latency_ms"],d["roofline"]["frac"]))
print({k:round(v["ms_per_launch"],2) for k,v in d["stages"].items()})
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r_new.json 2> gpurun_out/r_new.err && show gpurun_out/r_new.json
FRCS_BLOCK_SCATTER=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r_bs.json 2> gpurun_out/r_bs.err && show gpurun_out/r_bs.json
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_launches.csv $SMALL > gpurun_out/r_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r_launches.csv | grep -E "block_sort|plan|digits|scatter|ntt_chunk"
