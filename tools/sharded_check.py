#!/usr/bin/env python3
"""NCCL transport of the coset-sharded proof, under torchrun on a multi-GPU box:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_check.py [--num-tx T]

Every rank proves the same batch together with its peers (csg_dist_init over NCCL) and, alone, on its own GPU; the two
proofs must be byte-identical on every rank for several AIRs and sizes.  Prints one JSON line with the timings."""
import argparse
import hashlib
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-tx", type=int, default=64)
    args = ap.parse_args()
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch
    import torch.distributed as dist
    import certificate_stark_b200 as csg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("transaction", csg.AIR_TRANSACTION, *csg.TransactionBatch(seed=9, num_tx=args.num_tx).transaction_trace(), csg.ProofOptions()),
             ("transaction_sha3", csg.AIR_TRANSACTION, *csg.TransactionBatch(seed=2, num_tx=2).transaction_trace(), csg.ProofOptions(hash_fn=csg.HASH_SHA3_256)),
             ("merkle_update", csg.AIR_MERKLE_UPDATE, *csg.TransactionBatch(seed=3, num_tx=4).merkle_update_trace(), csg.ProofOptions()),
             ("schnorr", csg.AIR_SCHNORR, *csg.SignatureBatch(seed=5, num_sig=4).schnorr_trace(), csg.ProofOptions()),
             ("range", csg.AIR_RANGE, *csg.build_range_trace(987654321), csg.ProofOptions())]
    single, sharded = csg.Context(local), csg.Context(local)
    sharded.dist_init_torch()
    report = {"world": world, "transport": "nccl", "cases": {}}
    for name, air, trace, pub, opt in cases:
        want = single.prove(air, trace, pub, opt)
        t1 = single.timings()["total"]
        got = sharded.prove(air, trace, pub, opt)
        got = sharded.prove(air, trace, pub, opt)
        t = sharded.timings()
        ok = [None] * world
        dist.all_gather_object(ok, (got == want, hashlib.sha256(got).hexdigest()))
        report["cases"][name] = {"identical_on_every_rank": all(o[0] for o in ok) and len({o[1] for o in ok}) == 1, "proof_bytes": len(got),
                                 "single_ms": t1, "sharded_ms": t["total"], "comm_ms": t["comm"], "verify": csg.verify(air, pub, got)}
    if rank == 0:
        print(json.dumps(report), flush=True)
    bad = [k for k, v in report["cases"].items() if not v["identical_on_every_rank"] or v["verify"] != 0]
    single.close(); sharded.close()
    dist.destroy_process_group()
    if bad:
        raise SystemExit(f"sharded proof mismatch: {bad}")


if __name__ == "__main__":
    main()
