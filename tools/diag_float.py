#!/usr/bin/env python
"""Development diagnostic: precision = 'float' solver (C3f) against the reference generated with precision = 'float'
(float-rounded constants, double arithmetic -- SURVEY.md 8(a) note) and against the double solver; timing of both."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST
from oracle import refs

sol_f, spec_f, cfg = prebuilt.get('C3f_equMPC_ADMM')
sol_d, spec_d, _ = prebuilt.get('C3_equMPC_ADMM')
b = sysmodel.synthetic_batch(cfg['sys'], 16384, seed=41)
uf, kf, ef = refs.get('C3f_equMPC_ADMM')[0].solve_batch(b['x0'], b['xr'], b['ur'], threads=16)
ud, kd, ed = refs.get('C3_equMPC_ADMM')[0].solve_batch(b['x0'], b['xr'], b['ur'], threads=16)
for arith in (ARITH_FAST, ARITH_EXACT):
    u, k, e, info = sol_f.solve_batch(b['x0'], b['xr'], b['ur'], arith=arith)
    for nm_, (ur_, kr, er) in (('vs float ref', (uf, kf, ef)), ('vs double ref', (ud, kd, ed))):
        dk = np.abs(k - kr)
        rel = (np.abs(u - ur_) / np.maximum(1, np.abs(ur_))).max(axis=1)
        conv = er == 1
        print('float solver arith', arith, nm_, 'e_mismatch', int((e != er).sum()), 'dk hist', np.bincount(np.minimum(dk, 5)).tolist(),
              'max dk', int(dk.max()), 'u rel (dk<=1, conv) max %.2e' % rel[(dk <= 1) & conv].max(), 'p99.9 %.2e' % np.quantile(rel[conv], 0.999),
              'all max %.2e' % rel.max(), flush=True)
B = 1 << 18
bb = sysmodel.synthetic_batch(cfg['sys'], B, seed=100)
for name, sol in (('double', sol_d), ('float', sol_f)):
    best = min(sol.solve_batch(bb['x0'], bb['xr'], bb['ur'])[3]['kernel_ms'] for _ in range(3))
    a = sol.kernel_attributes()
    print(name, 'kernel_ms', best, 'Msolves/s', B / best / 1e3, a, flush=True)
