#!/usr/bin/env python
"""Development tool: output tiles per pass (NB) and prefetch distance (PF) of the HMPC engine on C5a."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.platforms import cuda_code
from spcies_b200.solver import CudaSolver
V = [(4, 8), (4, 10), (5, 8), (6, 6), (3, 10)]
spec, cfg = prebuilt.spec_for('C5a_HMPC_SADMM_split')
if sys.argv[1] == 'build':
    for nb, pf in V:
        cu, _ = cuda_code.emit(spec, save_name=f'V_hmpc_nb{nb}_pf{pf}')
        print(cuda_code.build(cu, extra_flags=(f'-DSPCIES_HMPC_MMA_NB={nb}', f'-DSPCIES_HMPC_MMA_PF={pf}')))
else:
    b = sysmodel.synthetic_batch(cfg['sys'], 1 << 15, seed=100)
    ref = None
    for nb, pf in V:
        sol = CudaSolver(os.path.join(ROOT, 'generated_solvers', f'V_hmpc_nb{nb}_pf{pf}.so'), spec)
        r = [sol.solve_batch(b['x0'], b['xr'], b['ur']) for _ in range(2)]
        best = min(x[3]['kernel_ms'] for x in r)
        same = True if ref is None else bool((r[0][1] == ref[1]).all() and (r[0][2] == ref[2]).all())
        ref = ref or r[0]
        print(nb, pf, best, (1 << 15) / best * 1e3, r[0][3]['block_threads'], r[0][3]['smem_bytes'], 'k,e same as first:', same, flush=True)
        sol.free()
