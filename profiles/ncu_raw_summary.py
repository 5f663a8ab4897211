#!/usr/bin/env python
"""Summarise an `ncu --set full` report: per-launch duration, DRAM bytes, occupancy, hit rates.
usage: ncu -i REPORT.ncu-rep --page raw --csv | python profiles/ncu_raw_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum']
idx = [h.index(w) for w in want if w in h]
print('\t'.join(h[i] for i in idx))
print('\t'.join(rows[1][i] for i in idx))
for r in rows[2:]:
    print('\t'.join(r[i][:70] for i in idx))
