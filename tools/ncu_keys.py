"""Key metrics of an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__waves_per_multiprocessor', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print('##', name[:90])
    for h, u, v in zip(hdr, units, vals):
        if h in keys or ('smsp__average_warp' in h and 'issue_stalled' in h and 'ratio' in h and float(v or 0) > 0.2):
            print(' ', h, u, v)
