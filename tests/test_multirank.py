"""world_size-2 gloo tests (CPU) of the multi-rank plumbing bench.py uses: per-rank batches, MAX-over-ranks timing,
rank-0-only reporting of the reference arm."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import os, sys, json
sys.path.insert(0, os.environ["CSG_ROOT"])
import torch, torch.distributed as dist
import numpy as np
import certificate_stark_b200 as csg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# every rank builds its own batch from its own seed (bench.py: seed = 1000 + rank): shards are independent proofs
trace, pub = csg.TransactionBatch(seed=1000 + rank, num_tx=1).transaction_trace()
digest = int(np.bitwise_xor.reduce(trace.ravel()) & np.uint64(0x7fffffffffffffff))
gathered = [None] * world
dist.all_gather_object(gathered, digest)
t = torch.tensor([10.0 + rank, 5.0 - rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, "distinct_batches": len(set(gathered)), "max": t.tolist()}))
dist.destroy_process_group()
"""


def run_torchrun(args, env_extra=None, timeout=600):
    env = dict(os.environ, CSG_ROOT=str(ROOT), OMP_NUM_THREADS="2", **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29653"] + args
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout, cwd=str(ROOT))


def test_two_ranks_shard_by_independent_batches(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    out = run_torchrun([str(w)])
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line == {"world": 2, "distinct_batches": 2, "max": [11.0, 5.0]}


def test_reference_arm_prints_once_under_torchrun():
    out = run_torchrun([str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--num-tx", "2"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["value"] > 0
    assert lines[0]["cpu_baseline"]["kind"] == "port" and lines[0]["e2e"]["h2d_bytes_per_step"] == 0
