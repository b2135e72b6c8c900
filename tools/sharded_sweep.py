#!/usr/bin/env python3
"""Kernel sweep at N GPUs (BASELINE.json configs[4] "... at 1/2/4/8 GPUs"): the LDE, row-hash + Merkle and FRI stages of ONE proof
sharded by coset over the N ranks, for state-transition traces of 2^13..2^21 rows x 94 columns (2^16..2^24 LDE rows), from the
stage timings of the sharded proof (CUDA events on each rank's proving stream, max over ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29620 tools/sharded_sweep.py [--out ...]

Algorithmic bytes as in tools/kernel_sweep.py (whole job, all ranks together); at N = 1 it runs unsharded (python tools/sharded_sweep.py)."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--max-tx", type=int, default=2048)
    args = ap.parse_args()
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch
    import certificate_stark_b200 as csg
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = ROOT / "MEASURED_PEAKS.json"
    hbm = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    ctx = csg.Context(local)
    if world > 1:
        ctx.dist_init_torch()
    w, b, rows = 94, 8, []
    ntx = 8
    while ntx <= args.max_tx:
        n = 1024 * ntx
        lde_rows = n * b
        batch = csg.TransactionBatch(seed=77, num_tx=ntx)          # same seed on every rank: the same batch
        pub = batch.public_inputs()
        ctx.set_air(csg.AIR_TRANSACTION, n, pub, csg.ProofOptions())
        ctx.build_transaction_trace(batch)
        ctx.prove_loaded()
        acc, reps = {}, 3
        for _ in range(reps):
            ctx.build_transaction_trace(batch)
            ctx.prove_loaded()
            t = ctx.timings()
            for k in ("lde", "commit_trace", "fri", "comm"):
                acc[k] = acc.get(k, 0.0) + t[k] / reps
        v = torch.tensor([acc["lde"], acc["commit_trace"], acc["fri"], acc["comm"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        lde_ms, commit_ms, fri_ms, comm_ms = [float(x) for x in v.cpu()]
        by = {"lde": w * (16 * n + 8 * n + 8 * lde_rows), "commit": lde_rows * (8 * w + 32) + 64 * (lde_rows - 1)}
        rec = {"n_gpus": world, "width": w, "log2_lde_rows": lde_rows.bit_length() - 1, "trace_rows": n,
               "lde": {"ms": round(lde_ms, 4), "algorithmic_gbs": round(by["lde"] / lde_ms / 1e6, 1), "frac_of_hbm_peak_all_gpus": round(by["lde"] / lde_ms / 1e6 / (hbm * world), 4)},
               "hash_rows_and_merkle": {"ms": round(commit_ms, 4), "algorithmic_gbs": round(by["commit"] / commit_ms / 1e6, 1),
                                        "frac_of_hbm_peak_all_gpus": round(by["commit"] / commit_ms / 1e6 / (hbm * world), 4)},
               "fri_all_layers_ms": round(fri_ms, 4), "exchange_ms_per_proof": round(comm_ms, 4)}
        rows.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        ntx *= 4
    if rank == 0 and args.out:
        Path(args.out).write_text(json.dumps({"hbm_peak_gbs_per_gpu": hbm, "rows": rows}, indent=1))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
