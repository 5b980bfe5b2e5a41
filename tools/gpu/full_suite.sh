#!/bin/bash
# the driver's round-end sequence: GPU tests, smoke, default bench, head bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/all_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --workload head > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "head bench exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench.json", "gpurun_out/bench_head.json"):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, j['metric'], round(j['value']), j['ms_per_step'], 'e2e', j['e2e'] and round(j['e2e']['value']), j['clocks'], 'roofline', j['roofline']['kernel'] if 'kernel' in j['roofline'] else '', round(j['roofline']['frac'] or 0,3), 'cpu', j['cpu_baseline'] and j['cpu_baseline']['value'])
PY
