#!/bin/bash
# One gpurun call: GPU tests, both bench arms, ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --logn 9 --no-cpu-baseline > gpurun_out/bench_f512.json 2> gpurun_out/bench_f512.err; echo "bench f512 rc=$?"
timeout 600 python tools/run_config3.py > gpurun_out/config3_n1.json 2>/dev/null; tail -1 gpurun_out/config3_n1.json
timeout 900 python tools/bench_split.py --kind 1 --logn 10 --steps 3 --warmup 1 > gpurun_out/split_sb_n1.json 2>/dev/null; tail -1 gpurun_out/split_sb_n1.json
SMALL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $SMALL > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:accum0_kernel -c 1 -o gpurun_out/prof_accum0 $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu accum0 rc=$?"
timeout 600 python tools/prof_witness.py 592 > gpurun_out/plain_w.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:witness_kernel\|r1cs_ -s 5 -c 5 -o gpurun_out/prof_witness python tools/prof_witness.py 592 > gpurun_out/ncu_full_w.log 2>&1
echo "ncu witness rc=$?"
for n in 592 4096 16384; do python tools/time_r1cs.py $n; done > gpurun_out/time_r1cs.jsonl 2>/dev/null; cat gpurun_out/time_r1cs.jsonl
ls -la gpurun_out | head -40
