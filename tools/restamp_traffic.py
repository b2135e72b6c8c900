#!/usr/bin/env python3
"""profiles/traffic.json was stamped with a fingerprint of ALL of csrc/ (`csrc_sha16`).  bench.py now compares a fingerprint of the
kernel sources only (`kernel_sha16`: csrc/ without the host driver, bench.HOST_DRIVER_FILES).  This script derives that stamp for an
existing capture WITHOUT touching its counters: it finds the capture's tree in git (the commit whose full fingerprint equals the
recorded csrc_sha16), computes the kernel fingerprint of that same tree and adds it.   python tools/restamp_traffic.py <commit>"""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

commit = sys.argv[1]


def read(name):
    return subprocess.run(["git", "show", f"{commit}:certificate_stark_b200/csrc/{name}"], cwd=ROOT, capture_output=True, check=True).stdout


path = ROOT / "profiles" / "traffic.json"
t = json.loads(path.read_text())
cap = t["_captured_at"]
full = bench.csrc_sha16(read=read, include_host_driver=True)
if full != cap["csrc_sha16"]:
    raise SystemExit(f"commit {commit} is not the tree the counters were captured from: {full} != {cap['csrc_sha16']}")
cap["kernel_sha16"] = bench.csrc_sha16(read=read)
cap["restamped"] = f"kernel_sha16 derived from the tree of commit {commit}, whose full fingerprint is the recorded csrc_sha16 (tools/restamp_traffic.py)"
path.write_text(json.dumps(t, indent=1) + "\n")
print("kernel_sha16", cap["kernel_sha16"], "| working tree:", bench.csrc_sha16(), "| fresh:", cap["kernel_sha16"] == bench.csrc_sha16())
