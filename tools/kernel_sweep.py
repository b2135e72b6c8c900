#!/usr/bin/env python3
"""Kernel sweep (BASELINE.json configs[4]): coset LDE + leaf hashing + Merkle commit + one FRI fold over 2^16..2^24 LDE rows
x trace width, on synthetic device-resident columns, reported against the measured HBM roofline.

  python tools/kernel_sweep.py [--out profiles/r1_kernel_sweep.json] [--max-log 24]

Algorithmic bytes: LDE 8n (trace in) + 8bn (LDE out) plus the iNTT's 16n per column; row hashing 8w+32 per
LDE row; Merkle 64 bytes per interior node; FRI fold 40 bytes per folded element.  Single GPU; under torchrun every rank can
run it on its own device (--device LOCAL_RANK): the sweep has no cross-GPU step."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import certificate_stark_b200 as csg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r1_kernel_sweep.json"))
    ap.add_argument("--min-log", type=int, default=16)
    ap.add_argument("--max-log", type=int, default=24)
    ap.add_argument("--blowup", type=int, default=8)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    peaks = ROOT / "MEASURED_PEAKS.json"
    hbm = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    rows = []
    with csg.Context(args.device) as ctx:
        for width in (2, 14, 56, 65, 94):
            for log_lde in range(args.min_log, args.max_log + 1, 2):
                lde_rows, b = 1 << log_lde, args.blowup
                n = lde_rows // b
                if width * lde_rows * 8 * 3 > 120e9:      # LDE + scratch + inputs must fit comfortably in 180 GB
                    continue
                for hash_fn, hname in ((csg.HASH_BLAKE3_256, "blake3"), (csg.HASH_SHA3_256, "sha3")):
                    if hash_fn == csg.HASH_SHA3_256 and width not in (14, 94):
                        continue
                    ms = ctx.sweep(width, n, b, hash_fn, iters=3)
                    by = {"lde": width * (16 * n + 8 * n + 8 * lde_rows), "hash_rows": lde_rows * (8 * width + 32),
                          "merkle": 64 * (lde_rows - 1), "fri_fold": 40 * (lde_rows // 4)}
                    rec = {"width": width, "log2_lde_rows": log_lde, "trace_rows": n, "blowup": b, "hash": hname}
                    for k, key in (("lde", "lde_ms"), ("hash_rows", "hash_rows_ms"), ("merkle", "merkle_ms"), ("fri_fold", "fri_fold_ms")):
                        gbs = by[k] / (ms[key] / 1e3) / 1e9 if ms[key] > 0 else None
                        rec[k] = {"ms": round(ms[key], 4), "algorithmic_gbs": round(gbs, 1) if gbs else None,
                                  "frac_of_hbm_peak": round(gbs / hbm, 4) if gbs else None}
                    rows.append(rec)
                    print(json.dumps(rec), flush=True)
    Path(args.out).write_text(json.dumps({"hbm_peak_gbs": hbm, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
