#!/bin/bash
# A/B timing under gpurun: bash tools/ab_bench.sh [lib ...] -- the default library first, then each alternative build (CSG_LIB)
for v in "" "$@"; do
  if [ -n "$v" ]; then export CSG_LIB=$PWD/$v; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
s=d['stage_ms']
print('${v:-default}', round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('lde','commit_trace','constraints','cons_rescue','cons_ecc_banks','cons_ecc_final','cons_rest','composition','ood_deep')})"
done
