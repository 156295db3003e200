#!/usr/bin/env python
"""Key metrics of the first kernel in an ncu report (development aid; writes nothing).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_active.avg']
for r in rows[2:]:
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-70s %s %s" % (k, r[i], units[i]))
    for i, k in enumerate(hdr):
        if 'issue_stalled' in k and 'per_issue_active' in k:
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if v >= 0.02:
                print("  stall %-40s %.3f" % (k.split('issue_stalled_')[1].split('_per_issue')[0], v))
    print()
