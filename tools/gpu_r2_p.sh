#!/bin/bash
# fused Y3 (mul_sub2) + stream run-length fit: parity first, then proofs/s and witnesses/s; piece-length sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_abi.py tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_witness.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log
show() { python - "$1" <<PY
import json,sys
d=json.load(open(sys.argv[1]))
w=d["witness"]
print(sys.argv[1], "value %.1f e2e %.1f ms/step %.1f lat %.2f roof %.3f | wit %.0f sat %.0f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["single_proof_latency_ms"],d["roofline"]["frac"],w["value"],w["satisfy_only"]))
print({k:round(v["ms_per_launch"],2) for k,v in d["stages"].items()})
PY
}
for lc in 64 96 128; do
  FRCS_LC0=$lc timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/p_lc$lc.json 2> gpurun_out/p_lc$lc.err || { echo "failed $lc"; tail -5 gpurun_out/p_lc$lc.err; continue; }
  show gpurun_out/p_lc$lc.json
done
FRCS_STREAM_NO_FIT=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/p_nofit.json 2> gpurun_out/p_nofit.err && show gpurun_out/p_nofit.json
