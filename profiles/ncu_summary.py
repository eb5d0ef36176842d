#!/usr/bin/env python
"""Summarise an .ncu-rep: duration, DRAM bytes, pipe utilisation, occupancy, stall reasons.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [extra-metric-substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:90])
    for i, h in enumerate(hdr):
        if h in KEYS or any(e in h for e in extra):
            print("  %-78s %-14s %s" % (h, units[i], r[i]))
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, n in sorted(stalls, reverse=True)[:8]:
        print("  stall %-30s %.3f warps/issue" % (n, v))
