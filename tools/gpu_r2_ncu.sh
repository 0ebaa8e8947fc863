#!/bin/bash
# ncu --set full captures of the kernels named in VERDICT r1 item 8 + this round's new kernels, on the small bench
# command (one group of 16 proofs per step), after the same command ran clean without ncu
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 3 --batch 16 --wbatch 592 --no-cpu-baseline --no-extra"
timeout 600 $SMALL > gpurun_out/n_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/n_plain.log; exit 1; }
cap() {  # name, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/n_$1 $SMALL > gpurun_out/n_$1.log 2>&1
  echo "ncu $1 rc=$?"
  ncu -i gpurun_out/n_$1.ncu-rep --page details > gpurun_out/n_$1_details.txt 2>&1
  grep -E "^  [a-zA-Z<]|Duration|DRAM Throughput|Compute \(SM\) Throughput|Registers Per|Achieved Occupancy|Issued Ipc Active|dram__bytes|L2 Hit" gpurun_out/n_$1_details.txt | cut -c1-150 | head -24
}
cap ntt "ntt_chunk_kernel" 12 3
cap accum0 "accum0_kernel" 9 3
cap reduce "bucket_reduce_kernel|narrow_reduce_kernel|reduce_channels_kernel" 9 3
cap r1cs "r1cs_stream_kernel|r1cs_bundle_kernel" 8 2
cap witness "witness_kernel" 4 1
rm -f gpurun_out/n_*.ncu-rep
