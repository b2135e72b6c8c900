#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py::test_device_witness_matches_host_builder tests/test_gpu_parity.py::test_example_facade -x -q > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_batch.log
timeout 300 python tools/batch_time.py 1024 > gpurun_out/batch_time.json 2> gpurun_out/batch_time.err; echo "batch_time rc=$?"; cat gpurun_out/batch_time.json; tail -3 gpurun_out/batch_time.err
