#!/bin/bash
# 2-GPU validation of the NCCL transport: byte identity on every rank, then the bench line with the sharded-proof leg
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_check.py --num-tx 64 > gpurun_out/sharded_check_n$N.json 2> gpurun_out/sharded_check_n$N.err; echo "check rc=$?"
tail -c 1500 gpurun_out/sharded_check_n$N.json; tail -3 gpurun_out/sharded_check_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d.get('sharded_proof'))[:1500])
PY
tail -3 gpurun_out/bench_n$N.err
