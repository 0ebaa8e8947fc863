#!/bin/bash
# plain atomics in the wide digit sort (default build) and the dedicated squaring variant on top of the fused Y3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_schoolbook.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log
show() { python - "$1" <<PY
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.1f e2e %.1f ms/step %.1f lat %.2f roof %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["single_proof_latency_ms"],d["roofline"]["frac"]))
print({k:round(v["ms_per_launch"],2) for k,v in d["stages"].items()})
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/q_plain.json 2> gpurun_out/q_plain.err && show gpurun_out/q_plain.json
if [ -f falcon_r1cs_b200/variants/sqr.so ]; then
FRCS_LIB=falcon_r1cs_b200/variants/sqr.so timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/q_sqr.json 2> gpurun_out/q_sqr.err && show gpurun_out/q_sqr.json
fi
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches.csv $SMALL > gpurun_out/q_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/q_launches.csv --group > gpurun_out/q_launches_summary.txt; grep -A45 "last proof group" gpurun_out/q_launches_summary.txt | grep -E "digits|scatter|accum|plan"
