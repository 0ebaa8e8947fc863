#!/bin/bash
# source-level stall sampling of r1cs_stream_kernel (592 signatures, verdict only)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:r1cs_stream_kernel -s 3 -c 1 -o gpurun_out/v_stream python tools/time_r1cs.py 592 10 3 > gpurun_out/v_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/v_stream.ncu-rep --page source --csv > gpurun_out/v_stream_source.csv 2> gpurun_out/v_src.err
ncu -i gpurun_out/v_stream.ncu-rep --page details > gpurun_out/v_stream_details.txt 2>&1
ls -la gpurun_out/v_stream*; head -3 gpurun_out/v_stream_source.csv | cut -c1-600
rm -f gpurun_out/v_stream.ncu-rep
