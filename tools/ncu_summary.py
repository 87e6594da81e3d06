#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct', 'gpu__dram_throughput', 'sm__warps_active.avg.pct', 'launch__registers_per_thread',
    'launch__occupancy', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct',
    'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_fp64', 'pipe_fp64', 'smsp__average_warp', 'issue_stalled',
    'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate', 'launch__grid_size', 'launch__block_size',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'achieved_occupancy', 'launch__waves', 'lts__t_sectors_srcunit_tex_op_read.sum',
    'lts__t_sectors_srcunit_tex_op_write.sum', 'l1tex__data_bank_conflicts', 'smsp__pcsamp_warps_issue_stalled']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[4][:60], 'block', r[7], 'grid', r[8])
    for h, u, v in zip(hdr, units, r):
        if any(p in h for p in pats):
            print(f'  {h} [{u}] = {v}')
