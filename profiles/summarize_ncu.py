#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the handful of numbers DESIGN.md / bench.py cite.
Usage: python profiles/summarize_ncu.py gpurun_out/prof_r1.ncu-rep > profiles/r1_ncu_summary.md"""
import csv, io, subprocess, sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__inst_executed.avg.per_cycle_elapsed', 'IPC per SM'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory wavefronts % of peak'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared-memory wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe active %'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('smsp__sass_inst_executed_op_local_ld.sum', 'local (spill) loads'),
    ('smsp__sass_inst_executed_op_local_st.sum', 'local (spill) stores'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard (smem)'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait (fixed latency)'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard'),
    ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'stall no_instruction'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe_throttle'),
    ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'stall branch_resolving'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected'),
]

def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]; units = rows[1]
    print(f'# ncu summary of `{path}`\n')
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"## {d.get('Kernel Name', '?')[:90]}\n")
        print('| metric | value | unit |\n|---|---|---|')
        for k, label in KEYS:
            if k in d:
                print(f'| {label} (`{k}`) | {d[k]} | {u[k]} |')
        print()

if __name__ == '__main__':
    main(sys.argv[1])
