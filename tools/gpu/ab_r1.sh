#!/bin/bash
# same-box A/B of the whole step: the round-1 tree (_r1/, built in place) against the current tree, alternating
mkdir -p gpurun_out
show() { python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=j['forward']['kernels']
print('$1', round(j['value']), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], ' '.join(f'{n[:5]}={v[\"ms_per_step\"]:.2f}' for n,v in k.items()))
"; }
for i in 1 2; do
  (cd _r1 && timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 30 2>/dev/null | show r1)
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-gpu-baseline --no-extra --steps 30 2>/dev/null | show now
done | tee gpurun_out/ab_r1.txt
