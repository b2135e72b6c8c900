#!/usr/bin/env python3
"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
  python tools/launches_by_kernel.py gpurun_out/launches_rN.csv > profiles/rN_launches_by_kernel.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[ik]).split("::")[-1]
    v = float(r[iv].replace(",", ""))
    v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(r[iu].rstrip("econd"), v / 1e6)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("kernel, launches (2 proofs: 1 warm-up + 1 timed), total ms under ncu (cold-cache, serialised), share")
for name, v in tot.most_common():
    print(f"{name:<60s} {cnt[name]:4d} {v:10.3f} ms {100 * v / total:5.1f}%")
