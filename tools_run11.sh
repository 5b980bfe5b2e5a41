#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py tests/test_encoder_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|Error|error|assert|rel err|parity.*mode|parity.*proc|parity.*fixture" gpurun_out/$b.log | tail -n 14
done
python tools/attn_trace.py 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'])
for k,v in j['forward']['kernels'].items(): print(k, round(v['ms_per_step'],3))
PY
