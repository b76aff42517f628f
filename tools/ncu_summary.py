#!/usr/bin/env python
"""Key metrics of one kernel from an .ncu-rep as a markdown table.

    python tools/ncu_summary.py REP.ncu-rep
"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg']


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U, V = rows[0], rows[1], rows[2]
    print("| metric | unit | value |\n|---|---|---|")
    for i, h in enumerate(H):
        stall = 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h
        if stall:
            try:
                if float(V[i]) < 0.15:
                    continue
            except ValueError:
                continue
        if h in WANT or stall:
            print(f"| {h} | {U[i]} | {V[i]} |")


if __name__ == "__main__":
    main()
