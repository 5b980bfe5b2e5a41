#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_head_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|Error|error|assert|parity|Mismatch|Max abs" gpurun_out/$b.log | tail -n 40
done
