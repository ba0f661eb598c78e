#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration on one B200 next to the reference C solver on one host core
(development tool; bench.py measures the headline configuration C2 under the full contract).

    python tools/config_times.py [scale]      # on the GPU box; writes gpurun_out/config_times.json

For each configuration: kernel-only and end-to-end (host buffers) solves/s of the FAST arithmetic, mean k, the e_flag = -1
count, parity of a sample against oracle/_ref, and the single-core rate of the instantiated reference template."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# save_name -> (GPU batch, CPU sample)
CONFIGS = {
    'C2_laxMPC_FISTA': (1 << 20, 1 << 15),
    'C3_equMPC_ADMM': (1 << 18, 2048),
    'C4_ellipMPC_ADMM_soc': (1 << 17, 512),
    'C5b_MPCT_EADMM': (1 << 17, 1024),
    'C5a_HMPC_SADMM_split': (1 << 15, 24),
    # the remaining solvers of SURVEY 8(a), at the reference tests' settings (N = 10, tol 1e-7, k_max 5000)
    'T_equMPC_FISTA': (1 << 17, 2048),
    'T_laxMPC_ADMM': (1 << 17, 2048),
    'T_ellipMPC_ADMM': (1 << 16, 2048),
    'T_HMPC_ADMM_split': (1 << 15, 512),
}


def main(scale):
    from spcies_b200 import prebuilt, sysmodel
    from oracle import refs
    out = {}
    for name, (B, S) in CONFIGS.items():
        B = max(256, int(B * scale))
        sol, spec, cfg = prebuilt.get(name)
        b = sysmodel.synthetic_batch(cfg['sys'], B, seed=100, with_r=sol.has_r)
        kw = dict(r=b['r']) if sol.has_r else {}
        sol.solve_batch(b['x0'][:4096], b['xr'][:4096], b['ur'][:4096], **({'r': b['r'][:4096]} if sol.has_r else {}))
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], **kw)
            dt = time.perf_counter() - t0
            if best is None or info['kernel_ms'] < best[0]['kernel_ms']:
                best = (info, dt)
        info, dt = best
        ref = refs.get(name)[0]
        kws = {'r': b['r'][:S]} if sol.has_r else {}
        t0 = time.perf_counter()
        ur_, kr, er = ref.solve_batch(b['x0'][:S], b['xr'][:S], b['ur'][:S], threads=1, **kws)
        cpu_dt = time.perf_counter() - t0
        same = k[:S] == kr
        rel = np.abs(u[:S] - ur_) / np.maximum(1.0, np.abs(ur_))
        out[name] = dict(
            batch=B, kernel_ms=info['kernel_ms'], kernel_solves_s=B / info['kernel_ms'] * 1e3, e2e_solves_s=B / dt,
            mean_k=float(k.mean()), n_not_converged=int((e == -1).sum()), launches=info['launches'], block=info['block_threads'],
            regs=info['regs_per_thread'], smem=info['smem_bytes'],
            cpu_1core_solves_s=S / cpu_dt, cpu_sample=S, speedup_vs_1core=B / info['kernel_ms'] * 1e3 / (S / cpu_dt),
            parity=dict(compared=S, e_flag_mismatch=int((e[:S] != er).sum()), max_abs_dk=int(np.abs(k[:S] - kr).max()),
                        u_rel_same_k=float(rel[same].max()) if same.any() else None))
        print(name, json.dumps(out[name]), flush=True)
        sol.free()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'config_times.json'), 'w'), indent=1)


if __name__ == '__main__':
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
