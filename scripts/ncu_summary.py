#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per kernel duration, DRAM traffic, throughput %, stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:100], ' grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:78s} {r[i]:>18s} {units[i]}")
    items = []
    for i, k in enumerate(hdr):
        if 'issue_stalled' in k and 'pcsamp' not in k and k.endswith('per_issue_active.ratio'):
            try: items.append((float(r[i]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError: pass
    print('  stalls (warps per issue-active):', ', '.join(f"{k}={v:.2f}" for v, k in sorted(items, reverse=True)[:7]))
