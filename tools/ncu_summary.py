#!/usr/bin/env python
"""Compact per-kernel table from an `ncu -i X.ncu-rep --page raw --csv` export (the --set full capture).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv > profiles/r01_ncu_full_summary_<tag>.csv
"""
import csv
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_wait"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) if c in hdr else None for c in COLS]
w = csv.writer(sys.stdout)
w.writerow(["#"] + [c + (f" [{units[i]}]" if i is not None and units[i] else "") for c, i in zip(COLS, idx)])
for n, r in enumerate(rows[2:]):
    out = []
    for c, i in zip(COLS, idx):
        v = r[i] if i is not None else ""
        if c == "Kernel Name":
            v = v.replace("aa::<unnamed>::", "").replace("aa::", "")
        out.append(v)
    w.writerow([n] + out)
