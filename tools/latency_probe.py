#!/usr/bin/env python
"""Development tool: single-instance call latency against the iteration count (slope = time per iteration, intercept = overhead)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ENGINE_MMA
name = sys.argv[1] if len(sys.argv) > 1 else 'C2_laxMPC_FISTA'
sol, spec, cfg = prebuilt.get(name)
b = sysmodel.synthetic_batch(cfg['sys'], 400, seed=7)
for eng, label in ((0, 'latency engine'), (ENGINE_MMA, 'MMA engine')):
    ks, ts, tk = [], [], []
    for i in range(400):
        x0, xr, ur = b['x0'][i:i + 1], b['xr'][i:i + 1], b['ur'][i:i + 1]
        best, bestk = 1e9, 1e9
        for rep in range(6):
            t0 = time.perf_counter()
            u, k, e, info = sol.solve_batch(x0, xr, ur, engine=eng)
            dt = time.perf_counter() - t0
            if rep:
                best = min(best, dt)
                bestk = min(bestk, info['kernel_ms'])
        ks.append(int(k[0])); ts.append(best * 1e6); tk.append(bestk * 1e3)
    ks, ts, tk = np.array(ks), np.array(ts), np.array(tk)
    sel = ks < 300
    A = np.vstack([ks[sel], np.ones(sel.sum())]).T
    (s1, c1), (s2, c2) = np.linalg.lstsq(A, ts[sel], rcond=None)[0], np.linalg.lstsq(A, tk[sel], rcond=None)[0]
    print(label, ': python call %.3f us/iteration + %.1f us;  launch-to-completion %.3f us/iteration + %.1f us;  median k %d' %
          (s1, c1, s2, c2, int(np.median(ks))), flush=True)
