#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attention_tc_kernel|gemm_tcgen05_kernel<256, 1|gemm_tcgen05_kernel<256, 0" -s 12 -c 3 -o gpurun_out/prof_v4 $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu2.log
