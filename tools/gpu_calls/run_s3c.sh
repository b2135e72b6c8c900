#!/bin/bash
# session-3 A/B: the linear rest of the transaction AIR as three kernels (default) against the single kernel (CSG_REST_WHOLE=1), and 5 / 6 resident CTAs for the parts
O=gpurun_out/s3c; mkdir -p $O
bash tools/ab_lib.sh build/p5/libcsg.so build/p6/libcsg.so > $O/ab_parts.txt 2>&1; cat $O/ab_parts.txt
CSG_REST_WHOLE=1 bash tools/ab_lib.sh > $O/ab_whole.txt 2>&1; cat $O/ab_whole.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_headline.py -m gpu -q -x -k "transaction or split or 1024 or facade" > $O/pytest_parts.log 2>&1; echo "pytest parts rc=$?"; tail -3 $O/pytest_parts.log
