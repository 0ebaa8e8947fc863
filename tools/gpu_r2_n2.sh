#!/bin/bash
# 2-GPU evidence: bench.py under torchrun (weak scaling, config 3, split schoolbook proof in `extra`), then the accum0 full capture on GPU 0
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/f_bench_n2.json 2> gpurun_out/f_bench_n2.err; echo "bench rc=$?"; tail -2 gpurun_out/f_bench_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/f_bench_n2.json"))
print("N=2 value %.1f e2e %.1f witness %.0f"%(d["value"],d["e2e"]["value"],d["witness"]["value"]))
print(json.dumps(d["extra"])[:1500])
PY
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:accum0_kernel.*FqParams.*16" -s 2 -c 1 -o gpurun_out/f_accum0 $SMALL > gpurun_out/f_ncu_accum0.log 2>&1
echo "ncu accum0 rc=$?"
ncu -i gpurun_out/f_accum0.ncu-rep --page details > gpurun_out/f_accum0_details.txt 2>&1
ncu -i gpurun_out/f_accum0.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/f_accum0_dram.csv 2>&1
cut -c1-50,200- gpurun_out/f_accum0_dram.csv | tail -3
rm -f gpurun_out/f_accum0.ncu-rep
