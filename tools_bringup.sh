#!/bin/bash
# Bring-up run: each GPU test file in its own process (a trapped kernel kills only that file), bounded by timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for f in tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py tests/test_encoder_gpu.py; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -n 40 gpurun_out/$b.log
done
