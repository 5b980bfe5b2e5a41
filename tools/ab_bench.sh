#!/bin/bash
# A/B on one box: tools/ab_bench.sh <other-lib.so> [bench args]   (alternates B, A, B, A)
OTHER=$1; shift
for i in 1 2; do
  for lib in "$OTHER" ""; do
    CBAS_B200_LIB=$lib python bench.py --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=j['forward']['kernels']
print('${lib:-current}'.split('/')[-1], round(j['value']), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], ' '.join(f'{n}={v[\"ms_per_step\"]:.3f}' for n,v in k.items()))
"
  done
done
