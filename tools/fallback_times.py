#!/usr/bin/env python
"""Development tool: the dense tensor-core fallback of the FISTA / ADMM solvers against their one-thread-per-instance kernel."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ENGINE_MMA, ENGINE_SCALAR
out = {}
for name in ('S4_laxMPC_FISTA', 'S4_equMPC_ADMM', 'S4_laxMPC_ADMM', 'S2_equMPC_ADMM', 'T_equMPC_ADMM_vrho'):
    sol, spec, cfg = prebuilt.get(name)
    B = 1 << 16
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=100)
    r = {}
    for eng, label in ((ENGINE_MMA, 'dense'), (ENGINE_SCALAR, 'scalar')):
        best = 1e30
        for _ in range(3):
            u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], engine=eng)
            best = min(best, info['kernel_ms'])
        r[label] = dict(kernel_ms=best, solves_per_s=B / best * 1e3, block=info['block_threads'], mean_k=float(k.mean()))
    r['speedup'] = r['scalar']['kernel_ms'] / r['dense']['kernel_ms']
    out[name] = r
    print(name, json.dumps(r), flush=True)
    sol.free()
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r2_fallback_times.json'), 'w'), indent=1)
