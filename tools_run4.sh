#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gemm_gpu.py tests/test_encoder_gpu.py; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|Error|error|assert" gpurun_out/$b.log | tail -n 20
done
