#!/bin/bash
# Fr multiplications kept on IMAD.WIDE (opaque M0) in ntt.cu / spmv.cu, narrow scatter from the scalars
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
show() { python - "$1" <<PY
import json,sys
d=json.load(open(sys.argv[1]))
w=d["witness"]
print(sys.argv[1], "value %.1f e2e %.1f ms/step %.1f lat %.2f roof %.3f | wit %.0f sat %.0f gen %.0f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["single_proof_latency_ms"],d["roofline"]["frac"],w["value"],w["satisfy_only"],w["generate_only"]))
print({k:round(v["ms_per_launch"],2) for k,v in d["stages"].items()})
PY
}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/s_new.json 2> gpurun_out/s_new.err && show gpurun_out/s_new.json
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_launches.csv $SMALL > gpurun_out/s_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/s_launches.csv | grep -E "accum0|accumN|digits|scatter_scalar"
