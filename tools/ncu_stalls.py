#!/usr/bin/env python3
"""Warp-stall breakdown per profiled launch from `ncu -i X.ncu-rep --page raw --csv` (ncu --set full): the top stall reasons by
smsp__average_warps_issue_stalled_*_per_issue_active (warps stalled on that reason per issued instruction)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names = rows[hdr]
cols = [(i, n) for i, n in enumerate(names) if "issue_stalled" in n and n.endswith("_per_issue_active.ratio") and "not_issued" not in n]
if not cols:
    cols = [(i, n) for i, n in enumerate(names) if "issue_stalled" in n]
ik = names.index("Kernel Name")
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    vals = []
    for i, n in cols:
        try:
            vals.append((float(r[i].replace(",", "")), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
    vals.sort(reverse=True)
    print(r[ik].split("(")[0].split("::")[-1][-40:], " | ".join(f"{n} {v:.2f}" for v, n in vals[:8]))
