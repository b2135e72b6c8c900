#!/bin/bash
python tools/small_latency.py --reps 200 > gpurun_out/small_sync.txt 2>&1
CSG_TIMER_NOSYNC=1 python tools/small_latency.py --reps 200 > gpurun_out/small_nosync.txt 2>&1
cat gpurun_out/small_sync.txt; echo ----; cat gpurun_out/small_nosync.txt
