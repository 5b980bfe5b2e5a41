#!/bin/bash
# isolated GEMM timings for the current build, a 4-stage build and the previous commit; then one ncu capture
mkdir -p gpurun_out
for lib in "" cbas_b200/_ab/libcbas_b200_WORKTREE_DGEMM_FORCE_STAGES4.so cbas_b200/_ab/libcbas_b200_HEAD.so; do
  echo "== lib: ${lib:-current}"
  CBAS_B200_LIB=$lib timeout 300 python tools/ln_gemm_bench.py 30 2>&1 | grep -E "^\{|Error|error" 
done | tee gpurun_out/ln_ab.txt
timeout 200 python tools/ln_gemm_bench.py 2 up down proj > gpurun_out/plain_ln.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel" -s 12 -c 36 -o gpurun_out/prof_ln_r02a python tools/ln_gemm_bench.py 2 up down proj > gpurun_out/ncu_ln.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_ln.log
