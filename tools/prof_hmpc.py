#!/usr/bin/env python
"""Development tool: one C5a (HMPC SADMM_split N = 50) batch, for ncu captures of hmpc_mma_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
sol, spec, cfg = prebuilt.get(sys.argv[2] if len(sys.argv) > 2 else 'C5a_HMPC_SADMM_split')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 13
b = sysmodel.synthetic_batch(cfg['sys'], B, seed=100)
for _ in range(2):
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'])
print(info)
