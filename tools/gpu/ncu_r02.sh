#!/bin/bash
# Round-2 evidence in one gpurun call: launch lists (encoder step, head pass) and --set full captures of one block's
# kernels and of one head chunk.  Every ncu command runs only after the same command has exited 0 without ncu.
mkdir -p gpurun_out
ENC="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-extra"
HEAD="python bench.py --workload head --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline"
$ENC > gpurun_out/plain_enc.log 2>&1 && $HEAD > gpurun_out/plain_head.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 180 --csv --log-file gpurun_out/launches_r02_final.csv $ENC > gpurun_out/ncu_a.log 2>&1; echo "enc launch list $?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 195 --csv --log-file gpurun_out/launches_r02_head.csv $HEAD > gpurun_out/ncu_b.log 2>&1; echo "head launch list $?"
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|attention_tc_kernel|layernorm_kernel" -s 200 -c 8 -f -o gpurun_out/prof_r02_block $ENC > gpurun_out/ncu_c.log 2>&1; echo "enc full $?"
ncu --set full --clock-control none --import-source on -k regex:"head_|gemm_tcgen05_kernel" -s 11 -c 7 -f -o gpurun_out/prof_r02_head $HEAD > gpurun_out/ncu_d.log 2>&1; echo "head full $?"
