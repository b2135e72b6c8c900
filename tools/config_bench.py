#!/usr/bin/env python3
"""Proving time of every BASELINE.json config on one B200 next to the CPU port (oracle), with proof-byte parity checked.

  python tools/config_bench.py [--out profiles/r1_configs.json] [--cpu-max-rows 131072]

configs[0]  state-transition example, smallest batches (1 and 4 transactions, default options)
configs[1]  benches/rescue.rs (chains 128..1024, blowup 4) and benches/merkle.rs (1/16/128 transactions)
configs[2]  benches/schnorr.rs (1/16/128 signatures) and benches/range.rs
configs[3]  benches/state_transition.rs shapes 16/128 and the 1024-transaction batch (bench.py's workload)
The CPU port is only run (and parity only checked) for traces up to --cpu-max-rows rows, to keep the run short."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import certificate_stark_b200 as csg  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def cases():
    seed = np.arange(42, 49, dtype=np.uint64)
    for tx in (1, 4):
        yield f"configs[0] state-transition {tx} tx", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=1, num_tx=tx).transaction_trace(), 8
    # the example binary's own defaults: 4 transactions, cubic extension (examples/state-transition.rs:58-66); and quadratic
    for ext in (3, 2):
        yield f"configs[0] state-transition 4 tx, field extension {ext}", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=1, num_tx=4).transaction_trace(), 8, ext
    for chain in (128, 256, 512, 1024):
        yield f"configs[1] rescue chain {chain}", csg.AIR_RESCUE, csg.build_rescue_trace(seed, chain), 4
    for tx in (1, 16, 128):
        yield f"configs[1] merkle-update {tx} tx", csg.AIR_MERKLE_UPDATE, csg.TransactionBatch(seed=2, num_tx=tx).merkle_update_trace(), 8
    for sig in (1, 16, 128):
        yield f"configs[2] schnorr {sig} signatures", csg.AIR_SCHNORR, csg.SignatureBatch(seed=3, num_sig=sig).schnorr_trace(), 8
    yield "configs[2] range 64-bit", csg.AIR_RANGE, csg.build_range_trace(2**63 - 1), 8
    for tx in (16, 128, 1024):
        yield f"configs[3] state-transition {tx} tx", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=4, num_tx=tx).transaction_trace(), 8
    for ext in (2, 3):
        yield f"configs[3] state-transition 1024 tx, field extension {ext}", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=4, num_tx=1024).transaction_trace(), 8, ext
    yield "configs[3] state-transition 128 tx, field extension 3", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=4, num_tx=128).transaction_trace(), 8, 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r1_configs.json"))
    ap.add_argument("--cpu-max-rows", type=int, default=131072)
    args = ap.parse_args()
    rows = []
    with csg.Context(0) as ctx:
        for name, air, (trace, pub), blowup, *rest in cases():
            ext = rest[0] if rest else 1
            opt = csg.ProofOptions(blowup_factor=blowup, field_extension=ext)
            ctx.set_air(air, trace.shape[1], pub, opt)
            ctx.load_trace(trace)
            proof = ctx.prove_loaded()
            reps = 5 if trace.shape[1] <= 131072 else 3
            ctx.timer_start()
            for _ in range(reps):
                ctx.reload_resident_trace()
                ctx.prove_loaded()
            gpu_ms = ctx.timer_stop() / reps
            t = ctx.timings()
            rec = {"config": name, "trace": f"{trace.shape[1]} x {trace.shape[0]}", "blowup": blowup, "gpu_ms": round(gpu_ms, 3), "proof_bytes": len(proof),
                   "kernel_launches": int(t["kernel_launches"])}
            if trace.shape[1] <= args.cpu_max_rows:
                t0 = time.perf_counter()
                want = O.prove(air, trace, pub, O.options(blowup=blowup, field_extension=ext))
                rec["cpu_port_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
                rec["proof_identical_to_oracle"] = bool(want == proof)
                rec["oracle_verifier_accepts"] = O.verify(air, pub, proof) == 0
                rec["speedup"] = round(rec["cpu_port_ms"] / gpu_ms, 1)
            rec["host_verifier_accepts"] = csg.verify(air, pub, proof) == 0
            rows.append(rec)
            print(json.dumps(rec), flush=True)
    Path(args.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
