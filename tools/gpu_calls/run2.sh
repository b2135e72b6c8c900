#!/bin/bash
# round-2 second GPU call: microbenchmark v2 (pipe-cycle metrics), the ncu counters of one proof -> profiles/traffic.json, one bench line
mkdir -p gpurun_out
./tools/microbench/int_pipes > gpurun_out/int_pipes2.txt 2>&1
ncu --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/int_pipes2_ncu.csv ./tools/microbench/int_pipes > gpurun_out/int_pipes2_ncu.log 2>&1
python bench.py --profile --steps 1 > gpurun_out/profile.json 2> gpurun_out/profile.err; echo "profile rc=$?"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/make_traffic.py gpurun_out/launches.csv gpurun_out/profile.json "${CSG_COMMIT:-uncommitted}" > gpurun_out/stage_counters.txt 2> gpurun_out/make_traffic.err; echo "traffic rc=$?"
cp profiles/traffic.json gpurun_out/traffic.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json | head -c 700
