#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 180 --csv --log-file gpurun_out/launches_v6.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel<256, 1, 2>" -s 14 -c 2 -o gpurun_out/prof_v6_up $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture exit $?"
