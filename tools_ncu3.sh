#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/gemm_one.py 3072 768 2 0 1 4"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -c 9 -o gpurun_out/prof_gemm_up $CMD > gpurun_out/ncu3.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu3.log
