#!/usr/bin/env python
"""Condense Nsight Compute output into the small text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv  profiles/rNN_launches.csv
    python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep  profiles/rNN_forward_full.md [kernel-regex]

`launches`: one line per kernel launch (name shortened, grid, block, device time) + per-kernel totals and shares.
`full`    : the metrics the roofline argument rests on (duration, pipe utilisation, issue, stalls, DRAM bytes, LSU
            wavefronts, occupancy) for every captured launch that matches the regex.
    python tools/ncu_summary.py traffic  gpurun_out/prof_X.ncu-rep  KERNEL CLIPS FRAMES STATES
`traffic` : dram__bytes_read.sum + dram__bytes_write.sum of the FIRST captured launch of KERNEL -> profiles/forward_traffic.json,
            keyed by kernel and stamped with the hash of that kernel's sources (viterbi_spl_b200.build.kernel_build_id), so
            that bench.py reports `roofline.traffic` only while the library it times IS the build that was captured.
    python tools/ncu_summary.py sass     profiles/rNN_sass_opcodes.txt
`sass`    : per-kernel counts of the Blackwell-specific opcodes (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
            UBLKCP = cp.async.bulk, UTMALDG = tensor-map TMA, FMNMX3, CREDUX, ...) from `cuobjdump -sass` of the built library.
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


def _num(x):
    return float(x.replace(',', ''))


def traffic(src, kernel, clips, frames, states):
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from viterbi_spl_b200 import build
    txt = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kernel in d.get('Kernel Name', ''):
            break
    else:
        raise SystemExit(f'{kernel} not in {src}')

    def in_bytes(key):
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[units[hdr.index(key)]]
        return _num(d[key]) * scale

    rd, wr = in_bytes('dram__bytes_read.sum'), in_bytes('dram__bytes_write.sum')
    dur_unit = units[hdr.index('gpu__time_duration.sum')]
    dur_ms = _num(d['gpu__time_duration.sum']) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(dur_unit, 1e-6)
    path = os.path.join(root, 'profiles', 'forward_traffic.json')
    try:
        with open(path) as fh:
            doc = json.load(fh)
    except Exception:
        doc = {}
    doc.setdefault('kernels', {})[kernel] = {
        'dram_bytes_per_launch': rd + wr, 'dram_bytes_read': rd, 'dram_bytes_write': wr,
        'clips': int(clips), 'frames': int(frames), 'states': int(states), 'duration_ms_under_ncu': dur_ms,
        'build_id': build.kernel_build_id(kernel), 'source': os.path.basename(src),
        'how': 'ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of one launch'}
    doc.pop('dram_bytes_per_launch', None), doc.pop('source', None), doc.pop('algorithmic_bytes_per_launch', None)
    with open(path, 'w') as fh:
        json.dump(doc, fh, indent=1)
    print(json.dumps(doc['kernels'][kernel]))


OPCODES = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'FMNMX3', 'FMNMX', 'FADD2',
           'FADD', 'FFMA2', 'FFMA', 'CREDUX', 'REDUX', 'LDGSTS', 'LDS', 'STS', 'LDG', 'STG', 'ELECT', 'UCGABAR', 'SHFL']


def sass(dst):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, 'viterbi_spl_b200', 'libvit_b200.so')
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    counts, cur = OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'^void ', '', name)[:110]
            counts[cur] = OrderedDict((o, 0) for o in OPCODES)
            continue
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            if op in counts[cur]:
                counts[cur][op] += 1
    with open(dst, 'w') as fh:
        fh.write('# cuobjdump -sass viterbi_spl_b200/libvit_b200.so (sm_100a): occurrences of selected opcodes per kernel\n')
        fh.write('# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (non-tensor bulk copy), UTMALDG = tensor-map TMA load,\n')
        fh.write('# SYNCS = mbarrier ops, FMNMX3 = 3-input fp32 max, CREDUX = redux.sync on fp32, LDGSTS = cp.async\n')
        for k, c in counts.items():
            nz = ', '.join(f'{o} {n}' for o, n in c.items() if n)
            fh.write(f'{k}\n    {nz}\n')
    print(dst, len(counts), 'kernels')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == 'traffic':
        traffic(*sys.argv[2:7])
    elif sys.argv[1] == 'sass':
        sass(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '.')
