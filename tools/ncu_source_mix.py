#!/usr/bin/env python
"""Dynamic instruction mix and shared-memory wavefronts per opcode from an ncu source-page CSV (development tool).
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > f.csv; python tools/ncu_source_mix.py f.csv [kernel_index]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
# several kernels may be concatenated: split at "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
for b in blocks:
    hdr = b['rows'][0]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = defaultdict(float); wf = defaultdict(float); wfi = defaultdict(float)
    tot = 0
    for r in b['rows'][1:]:
        if len(r) < len(hdr):
            continue
        src = r[ix['Source']].strip()
        toks = src.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        parts = op.split('.')
        key = parts[0]
        if key in ('LDS', 'STS', 'LDG', 'STG', 'LDTM', 'STTM', 'SHFL'):
            key = '.'.join(p for p in parts if p in (key, '64', '128', '32', 'U8', 'x2', 'x4') or p == parts[0] or p.startswith('x') or p.isdigit())
        n = float(r[ix['Instructions Executed']] or 0)
        ex[key] += n; tot += n
        wf[key] += float(r[ix['L1 Wavefronts Shared']] or 0)
        wfi[key] += float(r[ix['L1 Wavefronts Shared Ideal']] or 0)
    print('=====', b['name'][:90], 'total warp-instr %.3g' % tot)
    for k, v in sorted(ex.items(), key=lambda kv: -kv[1])[:28]:
        print('  %-14s %12.4g  %5.1f%%   smem wavefronts %10.4g (ideal %10.4g)  wf/inst %.2f' % (k, v, 100 * v / tot, wf[k], wfi[k], wf[k] / v if v else 0))
    print('  total wavefronts %.4g' % sum(wf.values()))
