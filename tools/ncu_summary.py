#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: the metrics B200_PROFILING.md asks for, per kernel."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==== ", r[idx['Kernel Name']][:90])
        for w in WANT:
            if w in idx:
                print(f"   {w} = {r[idx[w]]} {units[idx[w]]}")
        for h, i in idx.items():
            if 'issue_stalled' in h and 'average' in h:
                try:
                    if float(r[i]) > 0.1:
                        print(f"   stall {h.split('issue_stalled_')[1].split('_per_')[0]} = {float(r[i]):.2f}")
                except ValueError:
                    pass


if __name__ == "__main__":
    main(sys.argv[1])
