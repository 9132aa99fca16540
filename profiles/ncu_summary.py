#!/usr/bin/env python
"""Summarise an ncu report (--page raw --csv) into the handful of counters the design cites.
usage: ncu -i rep.ncu-rep --page raw --csv | python profiles/ncu_summary.py"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio']
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:70])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:78s} {r[i]:>18s} {units[i]}")
    st = sorted(((float(r[i].replace(',', '')) if r[i] else 0.0, h) for i, h in enumerate(hdr)
                 if 'issue_stalled' in h and h.endswith('_per_warp_active.pct')), reverse=True)[:6]
    for v, h2 in st:
        print(f"   top-stall {h2:68s} {v:10.2f}")
