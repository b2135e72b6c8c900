#!/bin/bash
# ncu captures of one proof of the bench workload (run under gpurun, after the plain run exits 0).  The .ncu-rep files stay on
# the box (tens of MB each); what comes back in gpurun_out/ is the launch list and the per-launch summaries of tools/ncu_summary.py.
set -x
TAG=${1:-r1}
D=/tmp/prof; mkdir -p $D gpurun_out
python bench.py --profile --steps 1 > gpurun_out/profile_plain_$TAG.json 2> gpurun_out/profile_plain_$TAG.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1
cap() {  # name, kernel regex, extra ncu args
    ncu --set full --clock-control none --import-source on -k "regex:$2" $3 -f -o $D/$1 python bench.py --profile --steps 1 > gpurun_out/ncu_$1_$TAG.log 2>&1
    ncu -i $D/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/ncu_$1_summary_$TAG.txt
}
cap cons "cons_" "-c 6"
cap ntt "ntt" "-s 8 -c 6"
[ -n "$SKIP_HASH" ] || cap hash "hash_rows|merkle_level" "-c 3"
ls -la $D gpurun_out | tail -20
