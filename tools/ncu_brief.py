#!/usr/bin/env python
"""Brief view of an .ncu-rep: the metrics the roofline discussion needs, one line per kernel launch.
    python tools/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:100])
    stalls = {}
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f'  {h:75s} {r[i]:>16s} {units[i]}')
        if 'pcsamp_warps_issue_stalled_' in h and not h.endswith('_not_issued'):
            try: stalls[h.split('stalled_')[1]] = float(r[i])
            except ValueError: pass
    tot = sum(stalls.values()) or 1
    print('  stalls: ' + ', '.join(f'{k} {100*v/tot:.0f}%' for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]))
