#!/bin/bash
# session-3: the coefficient exchange block by block (broadcasts) -- local transport on one GPU: the sharded parity tests
O=gpurun_out/s3d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_extension.py tests/test_gpu_headline.py -m gpu -q -x -k "sharded or world or ranks or split_exchange" > $O/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -3 $O/pytest_sharded.log
CSG_COEF_BLOCKWISE=1 timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > $O/pytest_sharded_ag.log 2>&1; echo "pytest all-gather rc=$?"; tail -2 $O/pytest_sharded_ag.log
