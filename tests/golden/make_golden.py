#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_vectors.json from the CPU oracle (run from the repo root).

The reference holds no golden vectors for this path (its tests are prove->verify round trips only)
and cannot be executed here (Rust + un-vendored winterfell fork), so these vectors pin the ORACLE against regressions and
pin the CUDA path to it; the Rescue known answers were derived independently from the reference's constants during the
survey and are kept separately in rescue_kat.json.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import certificate_stark_b200 as csg  # noqa: E402  (host-side witness builders only; no GPU needed)
from oracle import pyoracle as O  # noqa: E402


def main():
    out = {}
    seed = np.arange(42, 49, dtype=np.uint64)
    for chain, hash_fn in [(8, 2), (128, 2), (128, 3)]:
        trace, pub = csg.build_rescue_trace(seed, chain)
        proof, dbg = O.prove(O.AIR_RESCUE, trace, pub, O.options(blowup=4, hash_fn=hash_fn), want_debug=True)
        out[f"rescue_chain{chain}_hash{hash_fn}"] = {
            "pub": [int(v) for v in pub], "proof_len": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest(),
            "trace_root": bytes(dbg.trace_root).hex(), "constraint_root": bytes(dbg.constraint_root).hex(), "z": int(dbg.z)}
    trace, pub = csg.build_range_trace(123456789012345)
    proof = O.prove(O.AIR_RANGE, trace, pub, O.options(blowup=8))
    out["range_123456789012345"] = {"pub": [int(v) for v in pub], "proof_len": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest()}
    batch = csg.TransactionBatch(seed=1, num_tx=1)
    trace, pub = batch.transaction_trace()
    proof, dbg = O.prove(O.AIR_TRANSACTION, trace, pub, O.options(), want_debug=True)
    out["transaction_seed1_tx1"] = {"pub": [int(v) for v in pub], "trace_sha256": hashlib.sha256(trace.tobytes()).hexdigest(), "proof_len": len(proof),
                                    "proof_sha256": hashlib.sha256(proof).hexdigest(), "trace_root": bytes(dbg.trace_root).hex()}
    col = (np.arange(64, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) % np.uint64(O.P)
    lde = O.from_mont_fast(O.lde_column(O.to_mont_fast(col), 4))
    out["lde_64x4"] = {"column": [int(v) for v in col], "lde": [int(v) for v in lde]}
    (Path(__file__).parent / "oracle_vectors.json").write_text(json.dumps(out, indent=1))
    print("wrote oracle_vectors.json")


if __name__ == "__main__":
    main()
