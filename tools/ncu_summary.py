#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into profiles/: key counters and the SASS opcode mix per warp-stage.

    python tools/ncu_summary.py gpurun_out/prof_r1b.ncu-rep profiles/ncu_r1b_step_kernel.csv [warp_stages]
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    warp_stages = float(sys.argv[3]) if len(sys.argv) > 3 else 65536 * 10 / 32 * 4000
    raw = page(rep, 'raw')
    hdr, units, vals = raw[0], raw[1], raw[2]
    kname = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    lines = ['# ncu --set full --clock-control none, kernel: %s' % kname, 'metric,unit,value']
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            lines.append('%s,%s,%s' % (h, u, v))
    src = page(rep, 'source')
    h2 = src[1]
    iS, iE, iN = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in src[2:]:
        if len(r) <= iE:
            continue
        try:
            n, s = int(r[iE]), int(r[iN])
        except ValueError:
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS])
        op = m.group(2).split('.')[0] if m else r[iS][:10]
        ops[op] += n
        samp[op] += s
        tot += n
    lines.append('# SASS opcode mix: warp-instructions per (warp, stage); %.0f warp-stages in the launch' % warp_stages)
    lines.append('opcode,per_warp_stage,percent,stall_samples')
    lines.append('TOTAL,%.1f,100.0,%d' % (tot / warp_stages, sum(samp.values())))
    for op, n in ops.most_common(30):
        lines.append('%s,%.1f,%.1f,%d' % (op, n / warp_stages, 100.0 * n / tot, samp[op]))
    open(dst, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
