#!/usr/bin/env python
"""Concise summary of an ncu report: python tools/ncu_summary.py file.ncu-rep [launch index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2 + which]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w} [{units[i]}] = {r[i]}")
items = []
for i, h in enumerate(hdr):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            items.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
items.sort(reverse=True)
tot = sum(v for v, _ in items) or 1
print("stalls: " + ", ".join(f"{h} {v / tot * 100:.1f}%" for v, h in items[:7]))
