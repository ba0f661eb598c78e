#!/usr/bin/env python
"""Development tool: is a streamed engine bound by L2 bandwidth (per-SM rate rises with fewer CTAs) or by latency x ring depth?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
name = sys.argv[1] if len(sys.argv) > 1 else 'C5a_HMPC_SADMM_split'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 14
sol, spec, cfg = prebuilt.get(name)
b = sysmodel.synthetic_batch(cfg['sys'], B, seed=100)
for grid in (148, 111, 74, 37):
    best = 1e30
    for _ in range(2):
        u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], grid_blocks=grid)
        best = min(best, info['kernel_ms'])
    print(name, 'grid', grid, 'kernel_ms %.1f' % best, 'solves/s per CTA %.1f' % (B / best * 1e3 / grid), 'total %.0f' % (B / best * 1e3), flush=True)
