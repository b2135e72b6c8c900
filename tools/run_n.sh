#!/bin/bash
# N-GPU pass under gpurun --gpus N: NCCL transport check (bytes identical on every rank), the bench line (independent proofs + one
# sharded proof), and the sharded proof again with the coefficient exchange as block-wise broadcasts (CSG_COEF_BLOCKWISE=1) for comparison
N=${1:-2}; O=gpurun_out/n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29611 tools/sharded_check.py --num-tx 64 > $O/sharded_check.json 2> $O/sharded_check.err; echo "check rc=$?"
$TR --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
CSG_COEF_BLOCKWISE=1 $TR --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_blockwise.json 2> $O/bench_blockwise.err; echo "bench (block-wise broadcasts) rc=$?"
python - <<PY
import json
print(open('$O/sharded_check.json').read()[-400:])
for f in ('bench', 'bench_blockwise'):
    d=[json.loads(l) for l in open('$O/%s.json' % f) if l.startswith('{')][0]
    s=d.get('sharded_proof', {})
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), '| sharded ms', s.get('ms_per_proof'), 'speedup', s.get('speedup_vs_one_gpu'), 'comm', s.get('comm_ms_per_proof'), 'e2e', s.get('e2e_ms_per_proof'))
    print('   rank0 stages', {k: round(v,2) for k,v in s.get('stage_ms_rank0', {}).items() if k in ('lde','commit_trace','constraints','composition','ood_deep','fri','queries','comm')})
PY
