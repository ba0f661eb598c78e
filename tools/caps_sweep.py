#!/usr/bin/env python
"""Development tool: iteration-cap rounds of the FISTA tensor-core engine on C2 (1 Mi instances resident on the device)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spcies_b200 import prebuilt, sysmodel
sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
big = sysmodel.synthetic_batch(cfg['sys'], B, seed=100)
dev = torch.device('cuda', 0)
d = {k: torch.from_numpy(v).to(dev) for k, v in big.items()}
d_u = torch.empty((B, 2), dtype=torch.float64, device=dev)
d_k = torch.empty(B, dtype=torch.int32, device=dev)
d_e = torch.empty(B, dtype=torch.int32, device=dev)
k = None
for caps in [(96, 320), (), (16,), (24,), (32,), (48,), (64,), (96,), (24, 96), (32, 128), (32, 160), (48, 192), (64, 256), (16, 64, 256), (24, 96, 320),
             (32, 128, 400), (48, 160, 480)]:
    ms = []
    for i in range(4):
        info = sol.solve_batch_device(B, d['x0'].data_ptr(), d['xr'].data_ptr(), d['ur'].data_ptr(), d_u.data_ptr(), d_k.data_ptr(),
                                      d_e.data_ptr(), tail_mode=3 if caps else 1, tail_caps=caps)
        ms.append(info['kernel_ms'])
    kk = d_k.cpu().numpy()
    same = True if k is None else bool((kk == k).all())
    k = kk if k is None else k
    print(caps, 'kernel_ms %.3f' % min(ms[1:]), 'Msolves/s %.2f' % (B / min(ms[1:]) / 1e3), 'launches', info['launches'], 'parked', info['parked'],
          'same k', same, flush=True)
q = np.percentile(k, [50, 75, 90, 95, 99, 99.4])
print('k percentiles 50/75/90/95/99/99.4:', q, 'share of work below cap 32/64/96:', [float(np.minimum(k, c).sum() / k.sum()) for c in (32, 64, 96)])
