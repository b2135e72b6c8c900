#!/bin/bash
# ONE ncu --set full capture of every constraint and 1024-point NTT launch of the second (timed) proof of `bench.py --profile`,
# reduced on the box to the per-launch summary, the stall breakdown and the hottest instructions (the .ncu-rep stays there).
TAG=${1:-r1}
D=/tmp/prof; mkdir -p $D gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:cons_|ntt1024" -s 26 -c 26 -f -o $D/proof python bench.py --profile --steps 1 > gpurun_out/ncu_proof_$TAG.log 2>&1
ncu -i $D/proof.ncu-rep --page raw --csv 2>/dev/null > $D/raw.csv
python tools/ncu_summary.py < $D/raw.csv > gpurun_out/ncu_proof_summary_$TAG.txt
python tools/ncu_stalls.py < $D/raw.csv > gpurun_out/ncu_proof_stalls_$TAG.txt
for k in cons_low_kernel cons_ecc_low_kernel cons_item_kernel; do
  ncu -i $D/proof.ncu-rep --page source --csv -k "regex:$k" 2>/dev/null | python tools/ncu_hot_lines.py 30 > gpurun_out/hot_${k}_$TAG.txt 2>/dev/null
done
