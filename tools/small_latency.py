#!/usr/bin/env python3
"""Latency and throughput of the SMALL proofs (BASELINE.json configs[0]-[2]: the reference's own bench sizes), where a proof is a few
dozen short kernels and the Fiat-Shamir round trips to the host, not arithmetic, set the time.

  python tools/small_latency.py [--reps 200] [--threads 1,4,8] [--out profiles/r2_small_latency.json]

Per case: host wall time per proof of one context (trace resident), the sum of the device stage times, kernel launches, and the
aggregate proofs/s of T contexts driven by T host threads (each its own stream; the library releases the GIL in every call)."""
import argparse
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import certificate_stark_b200 as csg  # noqa: E402


def cases():
    seed = np.arange(42, 49, dtype=np.uint64)
    yield "range 64-bit", csg.AIR_RANGE, csg.build_range_trace(2**63 - 1), 8
    yield "rescue chain 128", csg.AIR_RESCUE, csg.build_rescue_trace(seed, 128), 4
    yield "rescue chain 1024", csg.AIR_RESCUE, csg.build_rescue_trace(seed, 1024), 4
    yield "merkle-update 1 tx", csg.AIR_MERKLE_UPDATE, csg.TransactionBatch(seed=2, num_tx=1).merkle_update_trace(), 8
    yield "schnorr 1 signature", csg.AIR_SCHNORR, csg.SignatureBatch(seed=3, num_sig=1).schnorr_trace(), 8
    yield "state-transition 1 tx", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=1, num_tx=1).transaction_trace(), 8
    yield "state-transition 4 tx", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=1, num_tx=4).transaction_trace(), 8
    yield "state-transition 16 tx", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=4, num_tx=16).transaction_trace(), 8


def run(ctx, reps):
    for _ in range(reps):
        ctx.reload_resident_trace()
        ctx.prove_loaded()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--threads", default="1,4,8")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    tlist = [int(x) for x in args.threads.split(",")]
    rows = []
    for name, air, (trace, pub), blowup in cases():
        opt = csg.ProofOptions(blowup_factor=blowup)
        ctxs = [csg.Context(0) for _ in range(max(tlist))]
        proofs = []
        for c in ctxs:
            c.set_air(air, trace.shape[1], pub, opt)
            c.load_trace(trace)
            proofs.append(c.prove_loaded())
            run(c, 5)
        assert all(p == proofs[0] for p in proofs)
        rec = {"case": name, "trace": f"{trace.shape[1]} x {trace.shape[0]}", "proof_bytes": len(proofs[0])}
        t = ctxs[0].timings()
        rec["kernel_launches"] = int(t["kernel_launches"])
        rec["stage_ms_sum"] = round(sum(t[k] for k in ("lde", "commit_trace", "constraints", "composition", "ood_deep", "fri", "queries")), 4)
        for T in tlist:
            th = [threading.Thread(target=run, args=(ctxs[i], args.reps)) for i in range(T)]
            t0 = time.perf_counter()
            for x in th:
                x.start()
            for x in th:
                x.join()
            dt = time.perf_counter() - t0
            if T == 1:
                rec["wall_ms_per_proof"] = round(dt / args.reps * 1e3, 4)
            rec[f"proofs_per_s_{T}ctx"] = round(T * args.reps / dt, 1)
        for c in ctxs:
            c.close()
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
