#!/usr/bin/env python
"""Development tool: one device-resident C3 (equMPC ADMM N = 20) batch, for ncu captures of admm_mma_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
sol, spec, cfg = prebuilt.get('C3_equMPC_ADMM')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
b = sysmodel.synthetic_batch(cfg['sys'], B, seed=100)
for _ in range(2):
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'])
print(info)
