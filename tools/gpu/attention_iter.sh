#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_kernels_gpu.py -q -m gpu -x -k "attention" -p no:cacheprovider 2>&1 | tail -5
timeout 300 python tools/attn_trace.py 2>&1 | tail -9
timeout 900 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['clocks'])
for k,v in j['forward']['kernels'].items(): print(k, round(v['ms_per_step'],3))
PY
