#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel launch + per-instruction SASS multiplicities / stall samples.
usage: tools/ncu_summary.py report.ncu-rep [--sass] [--iters N_WARP_ITERS] [--min 0.5]"""
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__cycles_elapsed.max',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
]


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==== ', r[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:84s} {r[i]:>18s} {units[i]}')
    if '--sass' not in sys.argv:
        return
    iters = float(sys.argv[sys.argv.index('--iters') + 1]) if '--iters' in sys.argv else None
    thr = float(sys.argv[sys.argv.index('--min') + 1]) if '--min' in sys.argv else 0.5
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--print-source', 'sass']))))
    hdr = rows[1]
    iA, iS, iI = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    body = []
    for r in rows[2:]:
        if len(r) < 10 or r[0] == 'Kernel Name':
            break
        if r[0] == 'Address':
            continue
        body.append(r)
    tot = sum(int(r[iI]) for r in body)
    tots = sum(int(r[iS]) for r in body)
    print(f'total warp-instructions {tot}, samples {tots}, static instructions {len(body)}')
    if iters is None:
        iters = max(int(r[iI]) for r in body[:200]) or 1
    print(f'per warp-iteration ({iters:.0f} iterations): {tot / iters:.1f} warp-instructions')
    for k, r in enumerate(body):
        c = int(r[iI])
        if c >= thr * iters:
            print(f'{k:5d} x{c / iters:6.2f} smp={100 * int(r[iS]) / max(tots, 1):5.2f}% {r[iA].strip()[:100]}')


if __name__ == '__main__':
    main()
