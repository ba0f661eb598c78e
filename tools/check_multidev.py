#!/usr/bin/env python
"""On a multi-GPU box: the C ABI's own sharding (spcies_batch_opts.n_devices, one host thread per device, contiguous
slices, no collective) gives the same results as one device."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
nd = sol.device_count()
b = sysmodel.synthetic_batch(cfg['sys'], 1 << 20, seed=5)
u1, k1, e1, i1 = sol.solve_batch(b['x0'], b['xr'], b['ur'], n_devices=1)
out = {'devices': nd, 'one_device_ms': i1['total_ms']}
for n in (2, 4, 8):
    if n > nd:
        break
    sol.solve_batch(b['x0'], b['xr'], b['ur'], n_devices=n)
    u, k, e, i = sol.solve_batch(b['x0'], b['xr'], b['ur'], n_devices=n)
    out[f'n{n}'] = dict(same=bool(np.array_equal(u, u1) and np.array_equal(k, k1) and np.array_equal(e, e1)), total_ms=i['total_ms'],
                        kernel_ms=i['kernel_ms'], sum_k=i['sum_k'] == i1['sum_k'])
print(json.dumps(out))
