#!/bin/bash
# launch list of two timed steps (run plain first; ONE ncu per gpurun call)   usage: ncu_launchlist.sh <tag>
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 180 --csv --log-file gpurun_out/launches_$1.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list exit $?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$1.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:50]].append(float(r[vi].replace(',','')))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])): print(k, len(v), round(sum(v)/len(v)/1000,1),'us', round(100*sum(v)/tot,1),'%')
PY
