#!/usr/bin/env python3
"""TransactionExample::new + prove with everything on the device (batch_gen.cu + witness_gen.cu + the prover), 1024 transfers:
   python tools/batch_time.py [num_tx]   -> one JSON line (device times from CUDA events; host times by wall clock)"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import certificate_stark_b200 as csg  # noqa: E402

num_tx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = {"num_tx": num_tx}
t0 = time.perf_counter()
host = csg.TransactionBatch(seed=1000, num_tx=num_tx)
out["host_builder_s"] = time.perf_counter() - t0
with csg.Context(0) as ctx:
    pub = ctx.build_batch(1000, num_tx)          # warm-up: tables, allocations
    for _ in range(3):
        t0 = time.perf_counter()
        pub = ctx.build_batch(1000, num_tx)
        wall = time.perf_counter() - t0
    out["device_builder_ms"] = ctx.timings()["batch_build"]
    out["device_builder_wall_ms"] = wall * 1e3      # includes the host plan (draws + tree shape)
    assert (pub == host.public_inputs()).all()
    assert (ctx.download_batch_records(num_tx) == host.packed_records()).all()
    ctx.set_air(csg.AIR_TRANSACTION, 1024 * num_tx, pub, csg.ProofOptions())
    for _ in range(2):
        ctx.timer_start()
        ctx.build_batch(1000, num_tx)
        ctx.build_transaction_trace_resident()
        proof = ctx.prove_loaded()
        out["batch_witness_proof_ms"] = ctx.timer_stop()
    out["witness_ms"] = ctx.timings()["h2d"]
    out["proof_verifies"] = csg.verify(csg.AIR_TRANSACTION, pub, proof, csg.ProofOptions()) == 0
print(json.dumps(out))
