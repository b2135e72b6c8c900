D=/tmp/prof; mkdir -p $D
ncu --set full --clock-control none --import-source on -k "regex:ntt1024_kernel" -s 2 -c 2 -f -o $D/nttA python bench.py --profile --steps 1 > gpurun_out/ncu_nttA.log 2>&1
ncu -i $D/nttA.ncu-rep --page raw --csv 2>/dev/null > $D/raw.csv
python tools/ncu_stalls.py < $D/raw.csv > gpurun_out/nttA_stalls.txt
python tools/ncu_summary.py < $D/raw.csv > gpurun_out/nttA_summary.txt
ncu -i $D/nttA.ncu-rep --page source --csv 2>/dev/null > $D/src.csv
head -3 $D/src.csv | cut -c1-1500 > gpurun_out/nttA_src_head.txt
python tools/ncu_hot_lines.py 60 < $D/src.csv > gpurun_out/nttA_hot.txt 2> gpurun_out/nttA_hot.err
grep -n "issue_stalled" $D/raw.csv | head -2 | cut -c1-300 >> gpurun_out/nttA_hot.err
