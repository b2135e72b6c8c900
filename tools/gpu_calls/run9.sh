#!/bin/bash
bash tools/ab_lib.sh build/nopf/libcsg.so > gpurun_out/ab_pf2.txt 2>&1; cat gpurun_out/ab_pf2.txt
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_headline.py > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu2.log
