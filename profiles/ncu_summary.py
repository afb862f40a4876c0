#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [--stalls]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    names = [r[hdr.index('Kernel Name')] for r in rows[2:]]
    print('kernels:', names)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print('%-70s %-10s %s' % (w, rows[1][i], [r[i] for r in rows[2:]]))
    if '--stalls' in sys.argv:
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                print('%-70s %s' % (h.replace('smsp__average_warps_issue_stalled_', ''), [r[i] for r in rows[2:]]))


if __name__ == '__main__':
    main()
