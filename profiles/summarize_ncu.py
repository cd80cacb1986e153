#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the handful of numbers DESIGN.md / bench.py cite.
Usage: python profiles/summarize_ncu.py gpurun_out/prof_r1.ncu-rep > profiles/r1_ncu_summary.md
       python profiles/summarize_ncu.py REPORT --metrics-json KERNEL_KEY BATCH SUMMARY_MD
           additionally merges {KERNEL_KEY: {dram_bytes_per_instance, binding, source, ...}} into profiles/r2_kernel_metrics.json,
           the file bench.py reads `roofline.traffic` / `roofline.binding` from (BATCH = instances in the profiled launch,
           SUMMARY_MD = the committed summary these numbers can be checked against)."""
import csv, io, json, os, subprocess, sys

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

PIPES = [  # candidates for "the resource that binds", each a %-of-peak metric
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory wavefronts (LSU data pipe)'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM'),
    ('lts__t_sectors.avg.pct_of_peak_sustained_elapsed', 'L2 sectors'),
]


def _num(v):
    return float(str(v).replace(',', ''))


def _bytes(v, unit):
    return _num(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]


def metrics_json(rows, hdr, units, key, batch, summary_md):
    want = key.split('<')[0]
    r = max((r for r in rows[2:] if want in dict(zip(hdr, r)).get('Kernel Name', '')),
            key=lambda r: _num(dict(zip(hdr, r)).get('gpu__time_duration.sum', 0) or 0))
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    dram = _bytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + _bytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
    pipes = [(lab, _num(d[k])) for k, lab in PIPES if k in d and d[k] not in ('', 'n/a')]
    top = max(pipes, key=lambda t: t[1])
    rec = {'dram_bytes_per_instance': dram / batch, 'batch_profiled': batch, 'source': summary_md,
           'kernel_ms_under_ncu': _num(d['gpu__time_duration.sum']) * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1, 'msecond': 1, 'second': 1e3, 's': 1e3}.get(u['gpu__time_duration.sum'], 1),
           'binding': {'resource': top[0], 'frac': top[1] / 100.0, 'source': summary_md, 'all': {lab: v / 100.0 for lab, v in pipes}},
           'ipc': _num(d.get('sm__inst_executed.avg.per_cycle_elapsed', 0) or 0),
           'registers_per_thread': int(_num(d.get('launch__registers_per_thread', 0) or 0)),
           'local_loads': _num(d.get('smsp__sass_inst_executed_op_local_ld.sum', 0) or 0),
           'bank_conflicts': _num(d.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 0) or 0),
           'smem_wavefronts': _num(d.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 0) or 0)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'r2_kernel_metrics.json')
    allm = json.load(open(path)) if os.path.exists(path) else {}
    allm[key] = rec
    with open(path, 'w') as f:
        json.dump(allm, f, indent=1, sort_keys=True)
    sys.stderr.write(f'merged {key} into {path}\n')


def main(path, extra=()):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]; units = rows[1]
    if extra and extra[0] == '--metrics-json':
        metrics_json(rows, hdr, units, extra[1], int(extra[2]), extra[3])
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
    main(sys.argv[1], sys.argv[2:])
