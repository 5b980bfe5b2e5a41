#!/bin/bash
# A/B of an environment knob on one box: tools/ab_env.sh VAR "v1 v2 ..." [bench args]   (alternates the values, twice)
VAR=$1; VALS=$2; shift 2
for i in 1 2; do
  for v in $VALS; do
    env "$VAR=$v" python bench.py --no-cpu-baseline --no-gpu-baseline --no-extra --no-e2e "$@" 2>/dev/null | grep -E "^\{" | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=j['forward']['kernels']
print('$VAR=$v', round(j['value']), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], ' '.join(n[:5]+'='+str(round(x['ms_per_step'],2)) for n,x in k.items()))
"
  done
done
