#!/usr/bin/env python
"""Summaries of an ncu report.

    python tools/ncu_summary.py file.ncu-rep [launch index]        # concise text
    python tools/ncu_summary.py file.ncu-rep --csv out.csv         # metric,unit,launch0,... (profiles/ format)
"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate\.pct|lts__t_sectors_srcunit_tex_op_read\.sum|"
    r"l1tex__t_sector_hit_rate\.pct|l1tex__throughput\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"l1tex__m_xbar2l1tex_read_bytes\.sum|l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_registers)|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
    r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__pipe_tensor_subpipe_dmma_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct_of_peak_sustained_active|"
    r"sm__cycles_elapsed\.max|sm__cycles_active\.avg|smsp__pcsamp_warps_issue_stalled_[a-z_]+)$"
)


def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, launches = load(rep)
    if len(sys.argv) > 3 and sys.argv[2] == "--csv":
        with open(sys.argv[3], "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
            for i, h in sorted(enumerate(hdr), key=lambda t: t[1]):
                if KEEP.match(h) and "not_issued" not in h:
                    wr.writerow([h, units[i]] + [r[i] for r in launches])
        return
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    r = launches[which]
    stalls = []
    for i, h in enumerate(hdr):
        if not KEEP.match(h) or "not_issued" in h:
            continue
        if "pcsamp_warps_issue_stalled" in h:
            try:
                stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        else:
            print(f"{h} [{units[i]}] = {r[i]}")
    stalls.sort(reverse=True)
    tot = sum(v for v, _ in stalls) or 1
    print("stalls: " + ", ".join(f"{h} {v / tot * 100:.1f}%" for v, h in stalls[:7]))


if __name__ == "__main__":
    main()
