#!/bin/bash
./tools/microbench/int_pipes > gpurun_out/int_pipes3.txt 2>&1
ncu -k regex:kd --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fp64.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/int_pipes3_ncu.csv ./tools/microbench/int_pipes > /dev/null 2>&1
grep -E "dfma|i2f|f2i|class 0 " gpurun_out/int_pipes3.txt
