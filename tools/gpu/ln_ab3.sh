#!/bin/bash
# A/B on one box: standalone LayerNorm kernels vs the fused path, full bench step, alternating
for i in 1 2; do
  for f in 0 1; do
    CBAS_B200_LN_FUSION=$f python bench.py --no-cpu-baseline --no-e2e --steps 30 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=j['forward']['kernels']
print('ln_fusion=$f', round(j['value']), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], ' '.join(f'{n[:5]}={v[\"ms_per_step\"]:.2f}' for n,v in k.items()))
"
  done
done
