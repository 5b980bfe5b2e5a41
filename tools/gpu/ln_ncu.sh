#!/bin/bash
# ncu --set full of the LN-fused GEMM variants in isolation: usage ln_ncu.sh <tag> <cases...>  (cases: up qkv proj down)
mkdir -p gpurun_out
TAG=$1; shift
timeout 200 python tools/ln_gemm_bench.py 2 "$@" > gpurun_out/plain_ln.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel" -c 40 -o gpurun_out/prof_ln_$TAG python tools/ln_gemm_bench.py 2 "$@" > gpurun_out/ncu_ln.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_ln.log
