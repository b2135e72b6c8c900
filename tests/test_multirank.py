"""world_size-2 gloo tests (CPU) of the multi-rank plumbing bench.py uses: per-rank batches, MAX-over-ranks timing,
rank-0-only reporting of the reference arm."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

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


# ---------------------------------------------------------------------------------------------- coset-sharded proof: host side
def test_shard_plans_partition_cosets_and_columns(csg):
    # csg_dist_plan is the geometry csg_set_air uses: over all ranks the LDE cosets, the ce cosets and the column blocks must
    # each be covered exactly once, in rank order (the all-gathers assume contiguous, rank-ordered slices)
    for width, blowup, ce in [(94, 8, 8), (65, 8, 4), (56, 8, 8), (14, 4, 4), (2, 8, 2), (58, 4, 4), (94, 16, 8), (94, 32, 8)]:
        for world in [1, 2, 4, 8, 16, 32]:
            if world > blowup:
                with pytest.raises(csg.CsgError):
                    csg.dist_plan(0, world, blowup, ce, width)
                continue
            plans = [csg.dist_plan(r, world, blowup, ce, width) for r in range(world)]
            cosets = [k for p in plans for k in range(p.first_coset, p.first_coset + p.num_cosets)]
            assert cosets == list(range(blowup))
            ce_cosets = [k for p in plans for k in range(p.first_ce_coset, p.first_ce_coset + p.num_ce_cosets)]
            assert ce_cosets == list(range(ce))
            for p in plans:     # a ce coset kc is LDE coset kc * (blowup / ce): it must lie inside the rank's block
                for kc in range(p.first_ce_coset, p.first_ce_coset + p.num_ce_cosets):
                    assert p.first_coset <= kc * (blowup // ce) < p.first_coset + p.num_cosets
            cols = [c for p in plans for c in range(p.first_column, p.first_column + p.num_columns)]
            assert cols == list(range(width))
            assert len({p.columns_per_rank for p in plans}) == 1 and plans[0].columns_per_rank * world >= width
            if world <= ce:
                assert len({p.num_ce_cosets for p in plans}) == 1        # equal slices in the composition all-gather
    with pytest.raises(csg.CsgError):
        csg.dist_plan(3, 2, 8, 8, 94)
    with pytest.raises(csg.CsgError):
        csg.dist_plan(0, 3, 8, 8, 94)


def test_trace_chunks_stay_inside_a_ranks_column_block(csg):
    # csg_dist_trace_chunks is the walk of stage 1 over a rank's column block (copy + extension chunk by chunk from host memory,
    # one chunk for a resident trace).  The last rank owns FEWER columns than columns_per_rank whenever world does not divide the
    # width (94 columns over 8 ranks: 7 x 12 + 10) and the trace buffers hold exactly `width` columns: a chunk that runs past the
    # block reads past the buffer (an 8-GPU bench of round 2 died of exactly that on rank 7).
    for width, blowup, ce in [(94, 8, 8), (65, 8, 4), (56, 8, 8), (14, 4, 4), (2, 8, 2), (58, 4, 4), (94, 32, 8)]:
        for world in [1, 2, 4, 8, 16, 32]:
            if world > blowup:
                continue
            for from_host in (False, True):
                covered = []
                for r in range(world):
                    p = csg.dist_plan(r, world, blowup, ce, width)
                    chunks = csg.dist_trace_chunks(r, world, blowup, ce, width, from_host)
                    assert sum(chunks) == p.num_columns and all(c >= 1 for c in chunks)
                    if not from_host:
                        assert chunks == ([p.num_columns] if p.num_columns else [])
                    elif p.num_columns:
                        assert chunks[0] <= 2          # the extension starts after the first column(s), not after the block
                    c = p.first_column
                    for sz in chunks:
                        covered += list(range(c, c + sz))
                        c += sz
                    assert c <= width
                assert covered == list(range(width))


ID_WORKER = r"""
import os, sys, json, hashlib
sys.path.insert(0, os.environ["CSG_ROOT"])
import torch.distributed as dist
import certificate_stark_b200 as csg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# what Context.dist_init_torch does before csg_dist_init: rank 0 creates the NCCL id, the default group hands it to everyone
box = [csg.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
plan = csg.dist_plan(rank, world, 8, 8, 94)
out = [None] * world
dist.all_gather_object(out, (hashlib.sha256(box[0]).hexdigest(), len(box[0]), plan.first_coset, plan.num_cosets, plan.first_column, plan.num_columns))
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
"""


def test_nccl_id_reaches_every_rank_and_plans_agree(tmp_path):
    w = tmp_path / "id_worker.py"
    w.write_text(ID_WORKER)
    out = run_torchrun([str(w)])
    assert out.returncode == 0, out.stderr[-2000:]
    rows = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("[")][-1])
    assert len(rows) == 2 and rows[0][0] == rows[1][0] and rows[0][1] == 128
    assert [r[2:] for r in rows] == [[0, 4, 0, 47], [4, 4, 47, 47]]
