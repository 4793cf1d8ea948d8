#!/usr/bin/env python
"""Print selected metrics of an `ncu --page raw --csv` dump, one line per metric, one column per launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[2:] or ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
    'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'issue_stalled']
for i, h in enumerate(hdr):
    if any(p in h for p in pats):
        print(h, f'[{units[i]}]', [d[i][:48] for d in data])
