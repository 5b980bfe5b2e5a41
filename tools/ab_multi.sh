#!/bin/bash
# compare several library builds on one box: tools/ab_multi.sh "<bench command printing one JSON line>" lib1 lib2 ... (use "" for the in-tree build)
CMD=$1; shift
for i in 1 2; do
  for lib in "$@"; do
    CBAS_B200_LIB=$lib bash -c "$CMD" 2>/dev/null | grep -E "^\{" | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l)
    if 'kernels_ms' in j: print('${lib:-current}'.split('/')[-1][-28:], j['tokens'], round(j['frames_per_s']), j['kernels_ms']['attention'])
    else: print('${lib:-current}'.split('/')[-1][-28:], round(j['value']), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], ' '.join(k[:4]+'='+str(round(v['ms_per_step'],2)) for k,v in j['forward']['kernels'].items()))
"
  done
done
