#!/bin/bash
# second half of the round-2 measurement pass, after the last source change: full GPU suite, bench lines, counter pass
O=gpurun_out/final; mkdir -p $O /tmp/prof
timeout 1700 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -9 $O/pytest_gpu.log
python bench.py --profile --steps 1 > $O/profile_plain.json 2> $O/profile_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum \
    --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --profile --steps 1 > $O/ncu_launches.log 2>&1
python tools/make_traffic.py $O/launches.csv $O/profile_plain.json final > $O/stage_counters.txt && cp profiles/traffic.json $O/traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python tools/small_latency.py --reps 200 --out $O/small_latency.json > $O/small_latency.log 2>&1
timeout 900 python tools/config_bench.py --out $O/configs.json --cpu-max-rows 16384 > $O/configs.log 2>&1; echo "configs rc=$?"
tail -c 900 $O/bench_n1.json
