#!/usr/bin/env python
"""Print the metrics we track from an ncu report: python tools/ncu_summary.py report.ncu-rep [kernel-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[0]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    name = r[H.index('Kernel Name')] if 'Kernel Name' in H else ''
    print("==", name[:70])
    for w in WANT:
        if w in H:
            print("  %-75s %s %s" % (w, r[H.index(w)], rows[1][H.index(w)]))
    for i, h in enumerate(H):
        if h.startswith('smsp__average_warp') or ('warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')):
            try:
                if float(r[i]) > 3.0:
                    print("  stall %-69s %s" % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), r[i]))
            except ValueError:
                pass
