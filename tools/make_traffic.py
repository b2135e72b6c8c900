#!/usr/bin/env python3
"""profiles/traffic.json from the ncu launch list of ONE proof, cut into the stages of the proof.

  python bench.py --profile --steps 1 > gpurun_out/profile.json                      # plain run first: must exit 0
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum --clock-control none -c 4000 --csv \
      --log-file gpurun_out/launches.csv python bench.py --profile --steps 1
  python tools/make_traffic.py gpurun_out/launches.csv gpurun_out/profile.json [commit] > profiles/rN_stage_counters.txt

The last `gpu_launches` launches of the list are the timed proof; `stage_launches` of the bench line (csg_timings) says how many of
them belong to each stage, in order.  Written for bench.py: per stage, warp instructions executed, DRAM bytes, ALU / FMA pipe
instructions, and the same for the dominant kernel family -- stamped with the fingerprint of the kernel sources they came from."""
import collections
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (csrc_sha16)

STAGES = ("lde", "commit_trace", "constraints", "composition", "ood_deep", "fri", "queries")


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    line = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")][-1]
    commit = sys.argv[3] if len(sys.argv) > 3 else "uncommitted"
    hdr = rows[0]
    iid, ik, im, iv, iu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        rec = launches.setdefault(int(r[iid]), {"kernel": re.sub(r"\(.*", "", r[ik]).split("::")[-1]})
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(r[iu].replace("second", "s").replace("nsecond", "ns"), v / 1e6)
        elif r[iu].lower().startswith(("kbyte", "mbyte", "gbyte")):
            v *= {"k": 1e3, "m": 1e6, "g": 1e9}[r[iu][0].lower()]
        rec[r[im]] = v
    ordered = [launches[k] for k in sorted(launches)]
    per_proof = int(line["gpu_launches"]) // max(int(line["steps"]), 1)
    proof = ordered[-per_proof:]
    sl = line["stage_launches"]
    assert sum(sl[s] for s in STAGES) == per_proof, (sl, per_proof)
    out = {"_captured_at": {"kernel_sha16": bench.csrc_sha16(), "csrc_sha16": bench.csrc_sha16(include_host_driver=True), "commit": commit,
                            "command": "ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_*.sum,smsp__inst_executed_pipe_{alu,fma}.sum "
                                       "--clock-control none python bench.py --profile --steps 1 (1024 tx); last proof of the run"},
           "stages": {}, "kernels": {}}
    at = 0
    total_ms = sum(l.get("gpu__time_duration.sum", 0.0) for l in proof)
    print(f"stage, launches, ms under ncu (cold cache, serialised), share, warp instructions, DRAM bytes, ALU-pipe / FMA-pipe instructions")
    for s in STAGES:
        part = proof[at:at + sl[s]]
        at += sl[s]
        agg = {"launches": len(part), "ms_ncu": sum(l.get("gpu__time_duration.sum", 0.0) for l in part),
               "warp_insts": sum(l.get("smsp__inst_executed.sum", 0.0) for l in part),
               "dram_bytes": sum(l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in part),
               "alu_insts": sum(l.get("smsp__inst_executed_pipe_alu.sum", 0.0) for l in part),
               "fma_insts": sum(l.get("smsp__inst_executed_pipe_fma.sum", 0.0) for l in part)}
        agg["pipes"] = {"alu_share_of_issued": agg["alu_insts"] / agg["warp_insts"] if agg["warp_insts"] else None,
                        "fma_share_of_issued": agg["fma_insts"] / agg["warp_insts"] if agg["warp_insts"] else None}
        agg["kernels"] = sorted({l["kernel"] for l in part})
        out["stages"][s] = agg
        print(f"{s:<13s} {len(part):4d} {agg['ms_ncu']:9.3f} ms {100 * agg['ms_ncu'] / total_ms:5.1f}%  {agg['warp_insts']:.4g}  {agg['dram_bytes']:.4g}  "
              f"{agg['alu_insts']:.4g} / {agg['fma_insts']:.4g}")
    byk = collections.defaultdict(lambda: collections.Counter())
    for l in proof:
        k = byk[l["kernel"]]
        k["launches"] += 1
        k["ms_ncu"] += l.get("gpu__time_duration.sum", 0.0)
        k["warp_insts"] += l.get("smsp__inst_executed.sum", 0.0)
        k["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        k["alu_insts"] += l.get("smsp__inst_executed_pipe_alu.sum", 0.0)
        k["fma_insts"] += l.get("smsp__inst_executed_pipe_fma.sum", 0.0)
    print("\nkernel, launches, ms under ncu, share, warp instructions, DRAM bytes, ALU / FMA pipe instructions")
    for name, k in sorted(byk.items(), key=lambda kv: -kv[1]["ms_ncu"]):
        out["kernels"][name] = dict(k)
        print(f"{name:<58s} {int(k['launches']):4d} {k['ms_ncu']:9.3f} ms {100 * k['ms_ncu'] / total_ms:5.1f}%  {k['warp_insts']:.4g}  {k['dram_bytes']:.4g}  "
              f"{k['alu_insts']:.4g} / {k['fma_insts']:.4g}")
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
