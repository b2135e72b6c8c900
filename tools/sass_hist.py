#!/usr/bin/env python3
"""Opcode histogram of one kernel from `cuobjdump -sass`: python tools/sass_hist.py <obj> <substring of the mangled name> [top]"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            hist[m.group(1).split(".")[0]] += 1
total = sum(hist.values())
for op, c in hist.most_common(top):
    print(f"{c:7d} {100 * c / total:5.1f}%  {op}")
print(f"{total:7d} total")
