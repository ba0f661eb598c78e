#!/usr/bin/env python
"""Development tool: one small FAST-arithmetic batch through every tensor-core engine (run under compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
SCALE = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ONLY = sys.argv[2].split(',') if len(sys.argv) > 2 else None      # optional: comma-separated solver names
for name, B in (('C2_laxMPC_FISTA', 700), ('T_equMPC_FISTA', 300), ('C3_equMPC_ADMM', 300), ('T_laxMPC_ADMM', 200), ('T_ellipMPC_ADMM', 120), ('C4_ellipMPC_ADMM_soc', 300),
                ('C5b_MPCT_EADMM', 40), ('C5a_HMPC_SADMM_split', 24), ('T_HMPC_ADMM_split', 100),
                # round 2: dense engine policies, the structured SOC engine, the dense fallbacks of FISTA / ADMM
                ('C6_MPCT_ADMM_cs', 100), ('C7_HMPC_ADMM', 60), ('C8_MPCT_ADMM_semiband', 100), ('T_ellipHMPC_ADMM', 40),
                ('S4_laxMPC_FISTA', 200), ('S4_equMPC_ADMM', 200), ('S2_laxMPC_ADMM', 200), ('T_equMPC_ADMM_vrho', 100)):
    if ONLY and name not in ONLY:
        continue
    sol, spec, cfg = prebuilt.get(name)
    B = max(9, int(B * SCALE))
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=9, with_r=sol.has_r)
    kw = dict(r=b['r']) if sol.has_r else {}
    if sol.nref == 3:
        z = lambda w: (b['x0'][:, :w] * 0.0)
        xr, ur = (b['xr'], z(sol.n) + 0.01, z(sol.n)), (b['ur'], z(sol.m), z(sol.m) + 0.02)
    else:
        xr, ur = b['xr'], b['ur']
    u, k, e, info = sol.solve_batch(b['x0'], xr, ur, **kw)
    extra = ''
    if 'FISTA' in name:
        u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], engine=2)       # (small batches default to the latency engine)
        u2, k2, e2, i2 = sol.solve_batch(b['x0'], b['xr'], b['ur'], tail_mode=3, tail_caps=(5, 20))
        extra = ' caps: same=%s launches=%d' % (bool((u == u2).all() and (k == k2).all()), i2['launches'])
        # the latency engine: one CTA per instance (launch), then the lingering server behind the single-instance symbol
        u3, k3, e3, i3 = sol.solve_batch(b['x0'][:12], b['xr'][:12], b['ur'][:12])
        for i in range(6):
            us, ks, es, _ = sol.solve(b['x0'][i], b['xr'][i], b['ur'][i])
            assert not name.startswith('C') or (ks == k3[i] and (us == u3[i]).all())   # (T_ solvers carry the debug payload: scalar kernel)
        extra += ' single: grid=%d' % i3['grid_blocks']
    print(name, 'ok', info['block_threads'], info['launches'], int(k.sum()), extra, flush=True)
    sol.free()
