#!/usr/bin/env python3
"""opcode histogram of the main loop body of every kernel in tools/microbench/int_pipes (cuobjdump -sass): the instructions between
the loop's first instruction and its backward branch, per trip"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, body = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); body[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
    if cur and m:
        body[cur].append((int(m.group(1), 16), m.group(2), m.group(3)))
for name, ins in body.items():
    back = [(a, o, r) for a, o, r in ins if o.startswith("BRA") and re.search(r"0x([0-9a-f]+)", r) and int(re.search(r"0x([0-9a-f]+)", r).group(1), 16) < a]
    if not back:
        continue
    a_end, _, r = max(back, key=lambda t: t[0] - int(re.search(r"0x([0-9a-f]+)", t[2]).group(1), 16))
    a_start = int(re.search(r"0x([0-9a-f]+)", r).group(1), 16)
    h = collections.Counter(o for a, o, _ in ins if a_start <= a <= a_end)
    tag = re.search(r"k[b]?ILi(\d+)", name)
    kind = ("butterfly " if "kbILi" in name else "class ") + (tag.group(1) if tag else name)
    print(f"{kind:14s} {sum(h.values()):5d} inst/trip-block: " + " ".join(f"{o}:{c}" for o, c in h.most_common(12)))
