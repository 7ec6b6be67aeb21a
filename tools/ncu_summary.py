#!/usr/bin/env python
"""One-line-per-launch summary of an .ncu-rep (raw page): duration, DRAM bytes, throughput %, pipes, occupancy, stalls."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "winst"), ("launch__grid_size", "grid"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ]
for r in data:
    name = r[idx["Kernel Name"]]
    parts = [name[:90]]
    for c, short in cols:
        if c in idx:
            v = r[idx[c]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{units[idx[c]] if short in ('dur', 'rd', 'wr') else ''}")
    print("  ".join(parts))
