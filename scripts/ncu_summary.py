#!/usr/bin/env python
"""Key counters of one kernel from `ncu -i X.ncu-rep --page raw --csv` output (stdin or file)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__cycles_elapsed.avg.per_second', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct']
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print('==', name)
    for h, u, v in zip(hdr, units, vals):
        m = re.match(r'smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio', h)
        if h in WANT:
            print(f'   {h} {u} {v}')
        elif m:
            try:
                if float(v) >= 0.05:
                    print(f'   stall {m.group(1)} {float(v):.2f}')
            except ValueError:
                pass
