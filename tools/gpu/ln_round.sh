#!/bin/bash
# fused-LayerNorm check round: kernel tests, isolated GEMM timings, encoder parity, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ln_fused_gpu.py tests/test_gemm_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/t1.log 2>&1; echo "t1 exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/t1.log | tail -15
timeout 300 python tools/ln_gemm_bench.py 40 2>&1 | grep -E "^\{|rror" | tee gpurun_out/ln_ab3.txt
timeout 700 python -m pytest tests/test_encoder_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/t2.log 2>&1; echo "t2 exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/t2.log | tail -15
timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["ms_per_step"], j["clocks"], {k:round(v["ms_per_step"],3) for k,v in j["forward"]["kernels"].items()})
PY
