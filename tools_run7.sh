#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "all gpu tests exit $?"; tail -n 15 gpurun_out/all_gpu_tests.log
