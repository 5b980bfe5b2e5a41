#!/bin/bash
# launch list of one 1M-frame head pass (plain run first; ONE ncu per gpurun call)
mkdir -p gpurun_out
CMD="python bench.py --workload head --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_head.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 1500 --csv --log-file gpurun_out/launches_head.csv $CMD > gpurun_out/ncu_head.log 2>&1
echo "launch list exit $?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_head.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(',','')))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])): print(k, len(v), round(sum(v)/len(v)/1000,1),'us', round(100*sum(v)/tot,1),'%')
PY
