#!/usr/bin/env python3
"""Proving time with FieldExtension None / Quadratic / Cubic at one batch size: python tools/ext_time.py [num_tx]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import certificate_stark_b200 as csg
ntx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = csg.TransactionBatch(seed=5, num_tx=ntx); pub = b.public_inputs()
with csg.Context(0) as c:
    for ext in (1, 2, 3):
        c.set_air(csg.AIR_TRANSACTION, 1024 * ntx, pub, csg.ProofOptions(field_extension=ext))
        c.build_transaction_trace(b)
        p = c.prove_loaded()
        c.timer_start()
        for _ in range(3):
            c.build_transaction_trace(b)   # the resident trace is consumed by the witness builder path; rebuild (19 ms) outside the stage timings
            p = c.prove_loaded()
        ms = c.timer_stop() / 3
        t = c.timings()
        print(f"ext={ext} tx={ntx} proof {len(p)} B verify {csg.verify(csg.AIR_TRANSACTION, pub, p)}  witness+prove {ms:.1f} ms  stages:",
              {k: round(v, 2) for k, v in t.items() if k in ("lde", "commit_trace", "constraints", "composition", "ood_deep", "fri", "queries")}, flush=True)
