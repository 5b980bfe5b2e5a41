#!/bin/bash
mkdir -p gpurun_out
for lib in "" cbas_b200/_ab/libcbas_b200_WORKTREE_DLN_CONS_NO_C1.so "" cbas_b200/_ab/libcbas_b200_WORKTREE_DLN_CONS_NO_C1.so; do
  echo "== lib: ${lib:-current}"
  CBAS_B200_LIB=$lib timeout 300 python tools/ln_gemm_bench.py 40 up qkv 2>&1 | grep -E "^\{|Error|error"
done | tee gpurun_out/ln_ab2.txt
timeout 200 python tools/ln_gemm_bench.py 2 up > gpurun_out/plain_ln.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel" -s 4 -c 4 -o gpurun_out/prof_ln_r02b python tools/ln_gemm_bench.py 2 up > gpurun_out/ncu_ln.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_ln.log
