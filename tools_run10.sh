#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload head --steps 5 > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "head bench exit $?"; tail -3 gpurun_out/bench_head.err; tail -1 gpurun_out/bench_head.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.json
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'], j['cpu_baseline'], j['roofline'])
for k,v in j['forward']['kernels'].items(): print(k, round(v['ms_per_step'],3))
PY
