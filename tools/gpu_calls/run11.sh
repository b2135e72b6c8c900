#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_extension.py "tests/test_gpu_headline.py::test_headline_sharded_proof_is_the_single_gpu_proof" tests/test_gpu_parity.py -x -q > gpurun_out/pytest_sh.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_sh.log
python bench.py --profile --steps 5 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]
s=d['stage_ms']
print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in s.items()})"
