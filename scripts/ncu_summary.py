#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics we track per kernel.
usage: scripts/ncu_summary.py report.ncu-rep [> profiles/xxx.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_inst"), ("sm__inst_executed_pipe_alu.sum", "alu"), ("sm__inst_executed_pipe_fma.sum", "fma"),
        ("sm__inst_executed_pipe_lsu.sum", "lsu"), ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"), ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math_throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_not_selected"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar")]
ki = hdr.index("Kernel Name")
seen = {}
for r in rows[2:]:
    name = r[ki].split("(")[0]
    seen.setdefault(name, []).append(r)
for name, rs in seen.items():
    r = rs[len(rs) // 2]
    print("== %s  (%d captures, showing the middle one)" % (name, len(rs)))
    for m, short in want:
        if m in hdr:
            i = hdr.index(m)
            print("   %-18s %s %s" % (short, r[i], units[i]))
