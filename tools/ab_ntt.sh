#!/bin/bash
# A/B of the 1024-point NTT pass shapes under gpurun (CSG_NTT_SHAPE_A / _B: 8 = 2 CTAs x 8 lanes, 4 = 4 x 4, 5 = 5 x 4 per SM)
for cfg in "8 8" "4 4" "5 5" "8 5" "5 8"; do
  set -- $cfg
  CSG_NTT_SHAPE_A=$1 CSG_NTT_SHAPE_B=$2 python bench.py --profile --steps 5 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
s=d['stage_ms']
print('A=$1 B=$2', round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('lde','constraints','composition','ood_deep')})"
done
