#!/usr/bin/env python
"""One 8 Mi-instance host-buffer call (BASELINE.json configs[4] batch size) on one device: every instance solved once, statistics
consistent, a random subset against the reference C solver."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel
from oracle import refs
name = sys.argv[1] if len(sys.argv) > 1 else 'C2_laxMPC_FISTA'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 23
sol, spec, cfg = prebuilt.get(name)
b = sysmodel.synthetic_batch(cfg['sys'], B, seed=3, with_r=sol.has_r)
kw = dict(r=b['r']) if sol.has_r else {}
t0 = time.perf_counter()
u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], **kw)
dt = time.perf_counter() - t0
idx = np.random.default_rng(1).integers(0, B, 4000)
ur_, kr, er = refs.get(name)[0].solve_batch(b['x0'][idx], b['xr'][idx], b['ur'][idx], threads=16, **({'r': b['r'][idx]} if sol.has_r else {}))
same = (k[idx] == kr) & (er == 1)
print(json.dumps(dict(name=name, B=B, seconds=dt, kernel_ms=info['kernel_ms'], solves_s=B / info['kernel_ms'] * 1e3,
                      sum_k_ok=bool(info['sum_k'] == int(k.sum())), nc_ok=bool(info['n_not_converged'] == int((e == -1).sum())),
                      e_values=sorted(set(np.unique(e).tolist())), k_min=int(k.min()), k_max=int(k.max()), launches=info['launches'],
                      parked=info['parked'], e_mismatch=int((e[idx] != er).sum()), max_dk=int(np.abs(k[idx] - kr).max()),
                      u_rel=float((np.abs(u[idx] - ur_) / np.maximum(1, np.abs(ur_)))[same].max()))))
