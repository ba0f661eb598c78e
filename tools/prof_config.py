#!/usr/bin/env python
"""Development tool: two host-buffer batches of one prebuilt solver, for ncu captures of its engine.
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1 -o out python tools/prof_config.py <solver> [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
name = sys.argv[1]
sol, spec, cfg = prebuilt.get(name)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
b = sysmodel.synthetic_batch(cfg['sys'], B, seed=100, with_r=sol.has_r)
kw = dict(r=b['r']) if sol.has_r else {}
for _ in range(2):
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], **kw)
print(info)
