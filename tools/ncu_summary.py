#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU): headline metrics, stall reasons, opcode mix and the hot basic blocks.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--blocks]
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --json profiles/ncu_c3.json    (the figures bench.py reports under roofline.ncu)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__sass_average_branch_targets_threads_uniform.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_global_ld.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed']


def ncu(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def as_json(rep, out):
    """Pipe utilisation as ncu measured it on this capture: what bench.py prints next to the algorithmic roofline fraction."""
    import json
    import os
    rows = ncu(rep, 'raw')
    hdr, units, val = rows[0], rows[1], rows[2]
    get = lambda k: float(val[hdr.index(k)].replace(',', ''))
    unit = lambda k: units[hdr.index(k)]
    dur = get('gpu__time_duration.sum') * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6}.get(unit('gpu__time_duration.sum'), 1e-3)
    ffma = get('smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed')
    fmul = get('smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed')
    fadd = get('smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed')
    lanes = 148 * 128  # FP32 lanes of the chip
    def mb(k):
        return get(k) * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit(k), 1e-6)
    d = {
        "capture": os.path.basename(rep), "kernel": val[hdr.index('Kernel Name')], "kernel_ms_under_ncu": dur,
        "pipe_fma_cycles_active_pct": get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
        "issue_slots_busy_pct": get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        "executed_fp32_frac_of_peak": (2 * ffma + fmul + fadd) / (2 * lanes),
        "ffma_thread_inst_per_cycle": ffma, "ffma_peak_per_cycle": lanes,
        "active_threads_per_warp_inst": get('smsp__thread_inst_executed_per_inst_executed.ratio'),
        "dram_read_mb": mb('dram__bytes_read.sum'), "dram_write_mb": mb('dram__bytes_write.sum'),
        "registers_per_thread": get('launch__registers_per_thread'),
        "note": "ncu --set full --clock-control none of ONE launch (cold caches, serialised); executed_fp32_frac_of_peak = "
                "(2 FFMA + FMUL + FADD thread instructions per cycle) / (2 x 148 x 128)",
    }
    with open(out, 'w') as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d, indent=1))


def main():
    rep = sys.argv[1]
    if '--json' in sys.argv:
        return as_json(rep, sys.argv[sys.argv.index('--json') + 1])
    rows = ncu(rep, 'raw')
    hdr, units = rows[0], rows[1]
    for val in rows[2:]:
        name = val[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        print('kernel:', name)
        for i, h in enumerate(hdr):
            if h in KEYS or ('issue_stalled' in h and 'per_issue_active' in h):
                print(f'{h:90s} {val[i]:>22s} {units[i]}')
    rows = ncu(rep, 'source')
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix['Instructions Executed']]) for r in data)
    print(f'\nSASS instructions {len(data)}, warp instructions executed {tot}')
    op, opthr = collections.Counter(), collections.Counter()
    for r in data:
        s = r[ix['Source']].split()
        o = (s[0] if not s[0].startswith('@') else s[1]).split('.')[0]
        op[o] += int(r[ix['Instructions Executed']])
        opthr[o] += int(r[ix['Thread Instructions Executed']])
    print('opcode mix (share of warp instructions, average active threads):')
    for o, c in op.most_common(18):
        print(f'  {o:10s} {c / tot * 100:5.1f}%  {opthr[o] / max(c, 1):5.1f}')
    if '--blocks' in sys.argv:
        print('hot basic blocks (>0.25 % of the warp instructions):')
        i = 0
        while i < len(data):
            c = int(data[i][ix['Instructions Executed']])
            j = i
            while j < len(data) and int(data[j][ix['Instructions Executed']]) == c:
                j += 1
            share = c * (j - i) / tot * 100
            if share > 0.25:
                ops = collections.Counter()
                for r in data[i:j]:
                    s = r[ix['Source']].split()
                    ops[(s[0] if not s[0].startswith('@') else s[1]).split('.')[0]] += 1
                smp = sum(int(r[ix['# Samples']]) for r in data[i:j])
                print(f'  [{i},{j}) n={j - i} exec={c} share={share:.1f}% threads={data[i][ix["Avg. Threads Executed"]]} samples={smp} {ops.most_common(5)}')
            i = j


if __name__ == '__main__':
    main()
