#!/bin/bash
# session-3 first pass: full GPU suite on HEAD, the counter pass that stamps traffic.json with these sources, the bench line
O=gpurun_out/s3a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -12 $O/pytest_gpu.log
python bench.py --profile --steps 1 > $O/profile_plain.json 2> $O/profile_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum \
    --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --profile --steps 1 > $O/ncu_launches.log 2>&1
python tools/make_traffic.py $O/launches.csv $O/profile_plain.json s3a > $O/stage_counters.txt && cp profiles/traffic.json $O/traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 600 $O/bench_n1.json
