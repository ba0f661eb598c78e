"""Diagnostic: precision = 'float' (float constants, double arithmetic) against the float-generated reference, per arithmetic / engine."""
import numpy as np
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, ENGINE_MMA, ENGINE_SCALAR
from oracle import refs
sol, spec, cfg = prebuilt.get('C3f_equMPC_ADMM')
ref = refs.get('C3f_equMPC_ADMM')[0]
b = sysmodel.synthetic_batch(cfg['sys'], 2048, seed=41)
ur_, kr, er = ref.solve_batch(b['x0'], b['xr'], b['ur'], threads=16)
for name, kw in (('exact', dict(arith=ARITH_EXACT)), ('fast scalar', dict(arith=ARITH_FAST, engine=ENGINE_SCALAR)), ('fast mma', dict(arith=ARITH_FAST, engine=ENGINE_MMA))):
    try:
        u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], **kw)
        same = k == kr
        print(name, 'k equal %.4f' % same.mean(), 'e equal', (e == er).all(), 'max |du| (same k, conv) %.3e' % np.abs(u - ur_)[same & (er == 1)].max(), info['block_threads'])
    except Exception as ex:
        print(name, 'failed:', ex)
