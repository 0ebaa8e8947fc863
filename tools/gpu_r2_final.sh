#!/bin/bash
# Round-2 evidence, one gpurun call on one B200: GPU suite, smoke, both bench arms, ncu launch list + full capture of the
# dominant kernel (each ncu pass only after the same command ran clean without it)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/f_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
timeout 1500 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/f_bench_n1.json"))
print("value %.1f e2e %.1f proofs/s  ms/step %.1f launches %d roof %.3f lat %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],d["roofline"]["frac"],d["single_proof_latency_ms"]))
print({k:(round(v["ms_per_launch"],3), v["launches"]) for k,v in d["stages"].items()})
w=d["witness"]; print("witness gen+check %.0f  gen %.0f  sat %.0f /s"%(w["value"],w["generate_only"],w["satisfy_only"]))
print({k:(v.get("value"), v.get("seconds")) for k,v in d["extra"].items()}); print(d["cpu_baseline"]["value"])
PY
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
timeout 600 $SMALL > gpurun_out/f_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/f_plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv $SMALL > gpurun_out/f_ncu_list.log 2>&1
echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/f_launches.csv > gpurun_out/f_launches_summary.txt; head -30 gpurun_out/f_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:accum0_kernel.*FqParams.*16" -s 2 -c 1 -o gpurun_out/f_accum0 $SMALL > gpurun_out/f_ncu_accum0.log 2>&1
echo "ncu accum0 rc=$?"
ncu -i gpurun_out/f_accum0.ncu-rep --page details > gpurun_out/f_accum0_details.txt 2>&1
ncu -i gpurun_out/f_accum0.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/f_accum0_dram.csv 2>&1
cut -c1-50,200- gpurun_out/f_accum0_dram.csv | tail -5
rm -f gpurun_out/f_accum0.ncu-rep
