#!/bin/bash
# session-3 A/B: Rescue residuals + linear rest as one mixed grid (CSG_CONS_MIX), H2D chunk schedules, and parity of both
O=gpurun_out/s3b; mkdir -p $O
bash tools/ab_lib.sh > $O/ab_default.txt 2>&1; cat $O/ab_default.txt
CSG_CONS_MIX=1 bash tools/ab_lib.sh > $O/ab_mix.txt 2>&1; cat $O/ab_mix.txt
CSG_CONS_MIX=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x -k "transaction or schnorr or split" > $O/pytest_mix.log 2>&1; echo "pytest mix rc=$?"; tail -3 $O/pytest_mix.log
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q -x -k "representations or prefetched or batch_after or 1024" > $O/pytest_chunks.log 2>&1; echo "pytest chunks rc=$?"; tail -3 $O/pytest_chunks.log
python tools/e2e_chunks.py 1024 > $O/e2e_chunks.txt 2>&1; cat $O/e2e_chunks.txt
