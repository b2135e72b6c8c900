#!/usr/bin/env python3
"""One line per profiled launch from an `ncu --set full` report: duration, registers, grid, occupancy, issue slots, ALU / FMA pipe
utilisation, DRAM bytes and throughput, cache hit rates, warp instructions.

  ncu -i gpurun_out/prof_x.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/rN_ncu_x_summary.txt
"""
import csv
import sys

COLS = [("t", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("issue%", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fmaheavy%", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"),
        ("dramR", "dram__bytes_read.sum"), ("dramW", "dram__bytes_write.sum"), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct"), ("warp_insts", "smsp__inst_executed.sum"),
        ("local_st", "smsp__inst_executed_op_local_st.sum")]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        name = r[idx["Kernel Name"]].split("(")[0]          # function name with its template arguments, parameter list dropped
        parts = [name.split("::")[-1][-60:]]
        for label, metric in COLS:
            if metric in idx:
                parts.append(f"{label} {r[idx[metric]]} {units[idx[metric]]}".rstrip())
        print(" | ".join(parts))


if __name__ == "__main__":
    main()
