#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_vit_kernels_gpu.py tests/test_encoder_gpu.py; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$b.log
done
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
