#!/usr/bin/env python
"""Condense Nsight Compute output into the small text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv  profiles/rNN_launches.csv
    python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep  profiles/rNN_forward_full.md [kernel-regex]

`launches`: one line per kernel launch (name shortened, grid, block, device time) + per-kernel totals and shares.
`full`    : the metrics the roofline argument rests on (duration, pipe utilisation, issue, stalls, DRAM bytes, LSU
            wavefronts, occupancy) for every captured launch that matches the regex.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'\(.*$', '', name)
    name = re.sub(r'<.*$', '', name)
    return name.replace('void ', '').strip()[:80]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ix = {k: hdr.index(k) for k in ('ID', 'Kernel Name', 'Block Size', 'Grid Size', 'Metric Name', 'Metric Value')}
    out, tot = [], OrderedDict()
    for r in rows[1:]:
        if r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        k = short(r[ix['Kernel Name']])
        ns = float(r[ix['Metric Value']].replace(',', ''))
        out.append((r[ix['ID']], k, r[ix['Grid Size']], r[ix['Block Size']], ns))
        tot[k] = tot.get(k, 0.0) + ns
    total = sum(tot.values())
    with open(dst, 'w') as fh:
        fh.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: '
                 'compare SHARES, not absolutes)\n')
        fh.write('# per-kernel totals\nkernel,launches,total_us,share\n')
        for k, ns in sorted(tot.items(), key=lambda kv: -kv[1]):
            n = sum(1 for o in out if o[1] == k)
            fh.write(f'{k},{n},{ns / 1e3:.1f},{ns / total:.4f}\n')
        fh.write('# every launch\nid,kernel,grid,block,duration_us\n')
        for o in out:
            fh.write(f'{o[0]},{o[1]},"{o[2]}","{o[3]}",{o[4] / 1e3:.2f}\n')


KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
    'launch__cluster_max_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'sm__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
]


def full(src, dst, pattern='.'):
    txt = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as fh:
        fh.write(f'# ncu --set full --clock-control none ({src.split("/")[-1]}); values per launch\n')
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            if not re.search(pattern, d.get('Kernel Name', '')):
                continue
            fh.write(f'\n## {short(d["Kernel Name"])}  grid {d.get("Grid Size")} block {d.get("Block Size")}\n\n')
            fh.write('| metric | value | unit |\n|---|---|---|\n')
            for k in KEYS:
                if k in d and d[k] != '':
                    fh.write(f'| {k} | {d[k]} | {units[hdr.index(k)]} |\n')
            fh.write('\nstall reasons (warps stalled per issue-active cycle):\n\n| reason | ratio |\n|---|---|\n')
            st = [(k, float(d[k].replace(',', ''))) for k in hdr
                  if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and d[k]]
            for k, v in sorted(st, key=lambda kv: -kv[1]):
                if v > 0.005:
                    fh.write(f'| {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]} | {v:.3f} |\n')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '.')
