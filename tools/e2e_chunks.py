#!/usr/bin/env python3
"""End-to-end time of csg_prove_trace (page-locked and pageable host memory) for several H2D chunk schedules:
   python tools/e2e_chunks.py [num_tx]   -- one process per schedule (CSG_H2D_CHUNKS is read once per process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json, numpy as np
sys.path.insert(0, %r)
import certificate_stark_b200 as csg
num_tx = int(sys.argv[1]); n = num_tx * 1024
ctx = csg.Context(0); pinned = csg.HostBuffer(94, n)
_, pub = csg.TransactionBatch(seed=1000, num_tx=num_tx).transaction_trace(out=pinned.array)
opt = csg.ProofOptions(); ctx.set_air(csg.AIR_TRANSACTION, n, pub, opt)
ref = None
out = {}
pageable = np.array(pinned.array)
for name, ptr in (("pinned", pinned.ptr), ("pageable", pageable.ctypes.data)):
    for _ in range(3): p = ctx.prove_trace_ptr(ptr)
    ref = ref or p; assert p == ref
    ctx.timer_start()
    for _ in range(10): ctx.prove_trace_ptr(ptr)
    out[name] = round(ctx.timer_stop() / 10, 3)
t = ctx.timings(); out["lde"] = round(t["lde"], 3); out["h2d"] = round(t["h2d"], 3)
print(json.dumps(out))
""" % ROOT

num_tx = sys.argv[1] if len(sys.argv) > 1 else "1024"
for sched in ("8", "2,6,8", "1,3,4,8", "4", "2,6,8,16", "16"):
    env = dict(os.environ, CSG_H2D_CHUNKS=sched)
    r = subprocess.run([sys.executable, "-c", CHILD, num_tx], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    print(sched, line[0] if line else "FAILED " + r.stderr[-400:], flush=True)
