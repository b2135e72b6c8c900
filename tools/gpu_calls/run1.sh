#!/bin/bash
# round-2 first GPU call: integer-pipe microbenchmark (plain + ncu pipe counters), the GPU test suite, one bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt; free -g >> gpurun_out/smi.txt
./tools/microbench/int_pipes > gpurun_out/int_pipes.txt 2>&1
ncu --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_adu.sum,smsp__inst_executed_pipe_xu.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/int_pipes_ncu.csv ./tools/microbench/int_pipes > gpurun_out/int_pipes_ncu.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_n1.json
