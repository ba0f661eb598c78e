#!/usr/bin/env python
"""Development tool: one small FAST-arithmetic batch through every tensor-core engine (run under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
SCALE = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
for name, B in (('C2_laxMPC_FISTA', 700), ('T_equMPC_FISTA', 300), ('C3_equMPC_ADMM', 300), ('T_laxMPC_ADMM', 200), ('T_ellipMPC_ADMM', 120), ('C4_ellipMPC_ADMM_soc', 300),
                ('C5b_MPCT_EADMM', 40), ('C5a_HMPC_SADMM_split', 24), ('T_HMPC_ADMM_split', 100)):
    sol, spec, cfg = prebuilt.get(name)
    B = max(9, int(B * SCALE))
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=9, with_r=sol.has_r)
    kw = dict(r=b['r']) if sol.has_r else {}
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], **kw)
    extra = ''
    if 'FISTA' in name:
        u2, k2, e2, i2 = sol.solve_batch(b['x0'], b['xr'], b['ur'], tail_mode=3, tail_caps=(5, 20))
        extra = ' caps: same=%s launches=%d' % (bool((u == u2).all() and (k == k2).all()), i2['launches'])
    print(name, 'ok', info['block_threads'], info['launches'], int(k.sum()), extra, flush=True)
    sol.free()
