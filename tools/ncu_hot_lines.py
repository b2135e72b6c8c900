#!/usr/bin/env python3
"""Top instructions / source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (kernel built with -lineinfo)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if any("Samples" in c for c in r))
names = rows[hdr]
samp = next(i for i, c in enumerate(names) if "Samples" in c)
print("columns:", names[:12], "...", file=sys.stderr)
src_i = next((i for i, c in enumerate(names) if c == "Source"), 0)
stall_cols = [i for i, c in enumerate(names) if c.startswith("stall_") or "Stall" in c]
data = []
for r in rows[hdr + 1:]:
    if len(r) != len(names):
        continue
    try:
        s = int(float(r[samp] or 0))
    except ValueError:
        continue
    top = sorted(((int(float(r[i] or 0)), names[i]) for i in stall_cols if r[i] not in ("", "0")), reverse=True)[:2]
    data.append((s, r[src_i].strip()[:110], top))
tot = sum(d[0] for d in data) or 1
for s, src, top in sorted(data, key=lambda d: -d[0])[:int(sys.argv[1]) if len(sys.argv) > 1 else 25]:
    print(f"{100.0 * s / tot:5.1f}%  {src}   {top}")
