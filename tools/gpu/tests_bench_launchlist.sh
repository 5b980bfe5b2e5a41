#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_vit_kernels_gpu.py tests/test_encoder_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|Error|error|assert|rel err|parity.*mode|parity.*proc|parity.*fixture" gpurun_out/$b.log | tail -n 14
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'])
for k,v in j['forward']['kernels'].items(): print(k, round(v['ms_per_step'],3))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 180 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "launch list exit $?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_v7.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:50]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])): print(k, len(v), round(sum(v)/len(v)/1000,1),'us')
PY
