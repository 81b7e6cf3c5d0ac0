#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU needed).

    python scripts/ncu_summary.py launches <launches.csv>
    python scripts/ncu_summary.py full <prof.ncu-rep>
"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio',
        'l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio']


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, data = rows[h], rows[h + 1:]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) > vi:
            agg[r[ki][:80]].append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':82s} {'n':>4s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:82s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot * 100:6.1f}%")


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('====', r[ki][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:88s} {r[i]:>18s} {units[i]}")


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
