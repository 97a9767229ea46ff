#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python tools/profile_summary.py launches gpurun_out/launches.csv  profiles/rNN_launches.txt
  python tools/profile_summary.py full     gpurun_out/prof.ncu-rep  profiles/rNN_kernel.txt
  python tools/profile_summary.py sass     <mangled-kernel-substr>  profiles/rNN_sass.txt
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_static', 'sm__cycles_elapsed.avg',
    'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__sass_inst_executed_op_local_ld.sum',
    'sm__sass_inst_executed_op_local_st.sum',
]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith('=='))]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
        f.write('# source: %s\n# %8s %14s %7s  kernel\n' % (src, 'launches', 'total_ms', 'share'))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%10d %14.3f %6.1f%%  %s\n' % (a[0], a[1] / 1e6, 100 * a[1] / tot, k[:160]))
        f.write('# total %.3f ms over %d launches\n' % (tot / 1e6, sum(a[0] for a in agg.values())))


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write('# ncu --set full --clock-control none --import-source on; source: %s\n' % src)
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
            f.write('\n## %s\n' % name[:200])
            for i, h in enumerate(hdr):
                if h in KEEP or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
                    f.write('%-86s %-16s %s\n' % (h, units[i], vals[i]))


def sass(pattern, dst):
    lib = os.path.join(ROOT, 'sterne_b200', 'libsterne_b200.so')
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    blocks = re.split(r'(?=\n\s*Function : )', out)
    with open(dst, 'w') as f:
        for b in blocks:
            m = re.search(r'Function : (\S+)', b)
            if not m or pattern not in m.group(1):
                continue
            body = [re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l) for l in b.splitlines()
                    if not re.match(r'^\s*/\* 0x[0-9a-f]+ \*/\s*$', l)]
            ops = collections.Counter()
            for l in body:
                mm = re.search(r'/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
                if mm:
                    ops[mm.group(1).split('.')[0]] += 1
            f.write('# cuobjdump -sass %s\n# function %s\n# opcode histogram: %s\n' %
                    (os.path.relpath(lib, ROOT), m.group(1), dict(ops.most_common())))
            f.write('\n'.join(body) + '\n')


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'sass': sass}[sys.argv[1]](sys.argv[2], sys.argv[3])
