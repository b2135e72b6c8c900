#!/bin/bash
# A/B under gpurun: bash tools/ab_lib.sh [lib ...] -- the default library first, then each alternative build (CSG_LIB); stage times of `bench.py --profile`
for v in "" "$@"; do
  if [ -n "$v" ]; then export CSG_LIB=$PWD/$v; fi
  python bench.py --profile --steps 5 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
s=d['stage_ms']
print('${v:-default}', round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('lde','commit_trace','constraints','cons_rescue','cons_ecc_banks','cons_ecc_low','cons_ecc_final','cons_rest','composition','ood_deep','fri')})"
done
