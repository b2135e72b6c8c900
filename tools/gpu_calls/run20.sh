#!/bin/bash
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/small_latency.py --reps 200 --out gpurun_out/small_latency.json > gpurun_out/small_lazy.txt 2>&1; cut -c1-330 gpurun_out/small_lazy.txt
python tools/host_trace.py 2> gpurun_out/host_trace_all.txt; grep -n "====" -A15 gpurun_out/host_trace_all.txt | tail -50
bash tools/ab_lib.sh
