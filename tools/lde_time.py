#!/usr/bin/env python3
"""LDE / row hash / Merkle / FRI fold timings at one shape (csg_k_sweep): python tools/lde_time.py [width logn blowup]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import certificate_stark_b200 as csg
w, logn, b = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (94, 20, 8)))
with csg.Context(0) as c:
    print(w, logn, b, c.sweep(w, 1 << logn, b, iters=5))
