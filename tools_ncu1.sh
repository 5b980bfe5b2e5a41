#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 180 --csv --log-file gpurun_out/launches_v1.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|attention_kernel" -s 30 -c 8 -o gpurun_out/prof_v1 $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu2.log
