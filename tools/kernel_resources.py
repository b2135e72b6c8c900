#!/usr/bin/env python3
"""Registers, spill stack, shared memory and static SASS size of every kernel in the built objects (no GPU needed):
  python tools/kernel_resources.py > profiles/rN_kernel_resources.txt"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
objs = sorted((ROOT / "certificate_stark_b200" / "lib" / "obj").glob("*.o"))
print(f"{'kernel':<64s} {'regs':>5s} {'stack':>6s} {'smem':>7s} {'SASS':>7s}  object")
for o in objs:
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(o)], capture_output=True, text=True).stdout
    if "Function" not in res:
        continue
    sass = subprocess.run(["cuobjdump", "-sass", str(o)], capture_output=True, text=True).stdout
    size, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            size[cur] = 0
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            size[cur] += 1
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        name = m.group(1)
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::|csg::|\(.*", "", dem).replace("void ", "")
        print(f"{dem[:64]:<64s} {m.group(2):>5s} {m.group(3):>6s} {m.group(4):>7s} {size.get(name, 0):>7d}  {o.name}")
