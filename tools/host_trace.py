#!/usr/bin/env python3
"""Host-side timeline of a small proof (CSG_HOST_TRACE): python tools/host_trace.py 2> trace.txt -- the last proof of each case."""
import os
import sys
from pathlib import Path

import numpy as np

os.environ["CSG_HOST_TRACE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import certificate_stark_b200 as csg  # noqa: E402

seed = np.arange(42, 49, dtype=np.uint64)
cases = [("range 64-bit", csg.AIR_RANGE, csg.build_range_trace(2**63 - 1), 8),
         ("rescue chain 128", csg.AIR_RESCUE, csg.build_rescue_trace(seed, 128), 4),
         ("state-transition 1 tx", csg.AIR_TRANSACTION, csg.TransactionBatch(seed=1, num_tx=1).transaction_trace(), 8)]
with csg.Context(0) as ctx:
    for name, air, (trace, pub), blowup in cases:
        ctx.set_air(air, trace.shape[1], pub, csg.ProofOptions(blowup_factor=blowup))
        ctx.load_trace(trace)
        for i in range(6):
            if i == 5:
                print(f"==== {name}", file=sys.stderr, flush=True)
            ctx.reload_resident_trace()
            ctx.prove_loaded()
