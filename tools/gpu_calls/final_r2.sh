#!/bin/bash
# Round-2 measurement pass on ONE B200 (run under gpurun): everything profiles/README.md quotes for a single GPU.
# Plain runs first (each must exit 0), profiler passes after; nothing printed under ncu is a bench value.
O=gpurun_out/final; mkdir -p $O /tmp/prof
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt; nproc >> $O/smi.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -x -q > $O/pytest_part.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_part.log; tail -2 $O/pytest_part.log
# 1. the bench lines
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
# 2. one proof: plain, then the counter pass the stage table and traffic.json are cut from
python bench.py --profile --steps 1 > $O/profile_plain.json 2> $O/profile_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum \
    --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --profile --steps 1 > $O/ncu_launches.log 2>&1
# 3. ncu --set full of the constraint and NTT launches of the timed proof, and of the row hash / Merkle levels
ncu --set full --clock-control none --import-source on -k "regex:cons_|ntt1024" -s 26 -c 26 -f -o /tmp/prof/proof python bench.py --profile --steps 1 > $O/ncu_proof.log 2>&1
ncu -i /tmp/prof/proof.ncu-rep --page raw --csv 2>/dev/null > /tmp/prof/raw.csv
python tools/ncu_summary.py < /tmp/prof/raw.csv > $O/ncu_proof_summary.txt
python tools/ncu_stalls.py < /tmp/prof/raw.csv > $O/ncu_proof_stalls.txt
ncu --set full --clock-control none --import-source on -k "regex:hash_rows" -c 2 -f -o /tmp/prof/hash python bench.py --profile --steps 1 > $O/ncu_hash.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:merkle_level" -c 2 -f -o /tmp/prof/merkle python bench.py --profile --steps 1 >> $O/ncu_hash.log 2>&1
for r in hash merkle; do
  ncu -i /tmp/prof/$r.ncu-rep --page raw --csv 2>/dev/null > /tmp/prof/raw_$r.csv
  python tools/ncu_summary.py < /tmp/prof/raw_$r.csv >> $O/ncu_hash_summary.txt
  python tools/ncu_stalls.py < /tmp/prof/raw_$r.csv >> $O/ncu_hash_stalls.txt
done
# 4. the other configs, the kernel sweep, the small proofs
timeout 900 python tools/config_bench.py --out $O/configs.json --cpu-max-rows 131072 > $O/configs.log 2>&1; echo "configs rc=$?"
timeout 900 python tools/kernel_sweep.py --out $O/kernel_sweep.json > $O/kernel_sweep.log 2>&1; echo "sweep rc=$?"
python tools/small_latency.py --reps 200 --out $O/small_latency.json > $O/small_latency.log 2>&1
python tools/host_trace.py 2> /tmp/prof/host_trace_all.txt; python - <<'P' > gpurun_out/final/host_trace.txt
t = open('/tmp/prof/host_trace_all.txt').read()
for part in t.split('==== ')[1:]:
    i = part.index('[csg host trace]')
    print('====', part[:i].strip()); print(part[i:])
P
timeout 600 python tools/batch_time.py > $O/batch_time.json 2> $O/batch_time.err
ls -la $O; tail -c 600 $O/bench_n1.json
