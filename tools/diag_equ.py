#!/usr/bin/env python
"""Development diagnostic: worst instances of the MMA engine against the reference, with / without SPCIES_FISTA_MMA_PRESCALE."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.platforms import cuda_code
from spcies_b200.solver import CudaSolver

BASES = ['T_equMPC_FISTA', 'T_laxMPC_FISTA', 'C2_laxMPC_FISTA']
if sys.argv[1] == 'build':
    for base in BASES:
        spec, cfg = prebuilt.spec_for(base)
        for ps in (0, 1):
            cu, _ = cuda_code.emit(spec, save_name=f'V_{base}_ps{ps}')
            print(cuda_code.build(cu, extra_flags=(f'-DSPCIES_FISTA_MMA_PRESCALE={ps}',)))
else:
    from oracle import refs
    for base in BASES:
        spec, cfg = prebuilt.spec_for(base)
        for seed in (31, 2, 77):
            b = sysmodel.synthetic_batch(cfg['sys'], 40000, seed=seed)
            ur_, kr, er = refs.get(base)[0].solve_batch(b['x0'], b['xr'], b['ur'], threads=16)
            for ps in (0, 1):
                sol = CudaSolver(os.path.join(ROOT, 'generated_solvers', f'V_{base}_ps{ps}.so'), spec)
                for eng in (2, 1):
                    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], engine=eng)
                    rel = (np.abs(u - ur_) / np.maximum(1, np.abs(ur_))).max(axis=1)
                    same = k == kr
                    w = np.argsort(-np.where(same, rel, 0))[:3]
                    print(base, 'seed', seed, 'ps', ps, 'engine', eng, 'e_mismatch', int((e != er).sum()), 'ndk', int((~same).sum()),
                          'max|dk|', int(np.abs(k - kr).max()), 'worst', [(int(i), float('%.2e' % rel[i]), int(kr[i]), int(er[i])) for i in w], flush=True)
                sol.free()
