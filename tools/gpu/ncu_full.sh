#!/bin/bash
# --set full capture of one layer's worth of GEMM + attention launches inside the bench (ONE ncu per gpurun call)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|attention_tc_kernel|layernorm_kernel" -s 200 -c 8 -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu2.log
