"""Print selected metrics from an `ncu --page raw --csv` dump (one column per launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__inst_executed.sum', 'sm__inst_executed_pipe_', 'smsp__cycles_active.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct', 'smsp__warp_issue_stalled', 'smsp__inst_executed_op_local', 'sm__cycles_elapsed.max',
    'smsp__average_warp', 'launch__occupancy', 'smsp__pcsamp_warps_issue_stalled']
for i, n in enumerate(h):
    if any(n.startswith(p) for p in pats):
        vals = [r[i] for r in rows[2:]]
        if all(v in ('0', '', 'n/a') for v in vals): continue
        print(n, '[%s]' % units[i], vals)
