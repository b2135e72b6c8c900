#!/usr/bin/env python3
"""Markdown tables and profiles/traffic.json from the per-launch ncu summaries (tools/ncu_summary.py output) of one round.

  python tools/make_profile_tables.py r1 > /tmp/tables.md      # reads profiles/r1_ncu_*_summary.txt, writes profiles/traffic.json
"""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def parse(path):
    rows = []
    for line in Path(path).read_text().splitlines():
        parts = [p.strip() for p in line.split("|")]
        rec = {"kernel": parts[0]}
        for p in parts[1:]:
            m = re.match(r"(\S+) ([\d.]+) ?(\S*)", p)
            if m:
                v = float(m.group(2))
                unit = m.group(3)
                if unit.startswith("Gbyte"):
                    v *= 1e9
                elif unit.startswith("Mbyte"):
                    v *= 1e6
                elif unit.startswith("Kbyte"):
                    v *= 1e3
                if m.group(1) == "t" and unit.startswith("us"):
                    v /= 1e3
                rec[m.group(1)] = v
        rows.append(rec)
    return rows


def label(k):
    for pat, name in [("cons_low_kernel<0, 0", "cons_low<TX,0> Rescue residuals (even cosets)"), ("cons_ecc_low", "cons_ecc_low<TX> curve formulas (even cosets)"),
                      ("cons_ecc_merge", "cons_ecc_merge<TX> banks, all cosets"), ("cons_item_kernel<0, 2", "cons_item<TX,2> final point addition"),
                      ("cons_low_kernel<0, 3", "cons_low<TX,3> linear rest (even cosets)"), ("cons_final", "cons_final divisors + boundary"),
                      ("ntt1024_kernel<1", "ntt1024<staged> pass A"), ("ntt1024_kernel<0", "ntt1024<direct> pass B"), ("ntt_pass", "ntt_pass (generic)"),
                      ("hash_rows", "hash_rows<Blake3>"), ("merkle_level", "merkle_level<Blake3>")]:
        if pat in k:
            return name
    return k.split(" | ")[0][:60]


def table(rows):
    out = ["| kernel | ms | regs | grid | warps active | issue slots | ALU pipe | FMA pipe | DRAM read+write | DRAM throughput | L2 hit |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        out.append(f"| {label(r['kernel'])} | {r.get('t', 0):.2f} | {int(r.get('regs', 0))} | {int(r.get('grid', 0))} | {r.get('warps%', 0):.0f} % | {r.get('issue%', 0):.0f} % | "
                   f"{r.get('alu%', 0):.0f} % | {r.get('fma%', 0):.0f} % | {(r.get('dramR', 0) + r.get('dramW', 0)) / 1e9:.2f} GB | {r.get('dram%', 0):.1f} % | {r.get('l2hit%', 0):.0f} % |")
    return "\n".join(out)


P = ROOT / "profiles"
cons = parse(P / f"{tag}_ncu_cons_summary.txt")
ntt = parse(P / f"{tag}_ncu_ntt_summary.txt")
hsh = parse(P / f"{tag}_ncu_hash_summary.txt")
print("### constraint stage\n" + table(cons) + "\n\n### transforms\n" + table(ntt) + "\n\n### commitments\n" + table(hsh))


def find(rows, pat):
    return next((r for r in rows if pat in r["kernel"]), None)


traffic = {"_source": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, python bench.py --profile --steps 1 (1024 tx); "
                      f"profiles/{tag}_ncu_cons_summary.txt", "_pipes": {}}
for key, pat in [("cons_rescue", "cons_low_kernel<0, 0"), ("cons_ecc_final", "cons_item_kernel<0, 2"), ("cons_ecc_banks", "cons_ecc_low"), ("cons_ecc_low", "cons_ecc_low"),
                 ("cons_rest", "cons_low_kernel<0, 3")]:
    r = find(cons, pat)
    if r:
        traffic[key] = int(r.get("dramR", 0) + r.get("dramW", 0))
        traffic["_pipes"][key] = {"alu_pct": round(r.get("alu%", 0), 1), "fma_pct": round(r.get("fma%", 0), 1), "issue_slots_pct": round(r.get("issue%", 0), 1),
                                  "warps_active_pct": round(r.get("warps%", 0), 1), "source": "ncu (the largest kernel of the phase)"}
(P / "traffic.json").write_text(json.dumps(traffic, indent=1))
