#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gemm_gpu.py tests/test_encoder_gpu.py; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|parity|Error|error" gpurun_out/$b.log | tail -n 30
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
