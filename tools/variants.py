#!/usr/bin/env python
"""Kernel-variant harness (development tool): build the C2 laxMPC-FISTA solver under several sets of compile-time
switches and time each on one B200.

    python tools/variants.py build            # here (nvcc cross-compiles), writes generated_solvers/V_<name>.so
    python tools/variants.py run [B]          # on the GPU box: parity (EXACT vs oracle/_ref, 4096 instances) + timing

Every variant is a complete generated library, loaded through the same C ABI as the product solver."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    'default': [],
    'bulk384_ss': ['-DSPCIES_FISTA_MMA_BULK_SMEM=1'],
    'bulk512_ss': ['-DSPCIES_FISTA_MMA_BLOCK_BULK=512', '-DSPCIES_FISTA_MMA_BULK_SMEM=1'],
    'bulk256': ['-DSPCIES_FISTA_MMA_BLOCK_BULK=256'],
    'nomerge': ['-DSPCIES_FISTA_MMA_MERGE=0'],
    'scalar': ['-DSPCIES_FISTA_MMA=0'],
}
VARIANTS.update(json.loads(os.environ.get('SPCIES_VARIANTS', '{}')))
LIST = os.path.join(ROOT, 'generated_solvers', 'variants.json')


def build():
    from spcies_b200 import prebuilt
    from spcies_b200.platforms import cuda_code
    spec, cfg = prebuilt.spec_for('C2_laxMPC_FISTA')
    done = {}
    for name, flags in VARIANTS.items():
        save = 'V_' + name
        cu, _ = cuda_code.emit(spec, save_name=save)
        so = cuda_code.build(cu, extra_flags=tuple(flags))
        log = open(cu[:-3] + '.stamp').read()
        regs = [l.strip() for l in log.split('\n') if 'Used' in l or 'spill' in l]
        print(save, flags, regs[-2:])
        done[name] = flags
    json.dump(done, open(LIST, 'w'))


def run(B):
    import torch
    from spcies_b200 import prebuilt, sysmodel
    from spcies_b200.solver import CudaSolver, ARITH_EXACT, ARITH_FAST
    from oracle import refs
    spec, cfg = prebuilt.spec_for('C2_laxMPC_FISTA')
    ref = refs.get('C2_laxMPC_FISTA')[0]
    small = sysmodel.synthetic_batch(cfg['sys'], 4096, seed=5)
    ur_, kr, er = ref.solve_batch(small['x0'], small['xr'], small['ur'], threads=8)
    dev = torch.device('cuda', 0)
    big = sysmodel.synthetic_batch(cfg['sys'], B, seed=100)
    d = {k: torch.from_numpy(v).to(dev) for k, v in big.items()}
    d_u = torch.empty((B, 2), dtype=torch.float64, device=dev)
    d_k = torch.empty(B, dtype=torch.int32, device=dev)
    d_e = torch.empty(B, dtype=torch.int32, device=dev)
    out = {}
    only = os.environ.get('SPCIES_ONLY')
    for name in json.load(open(LIST)):
        if only and name != only:
            continue
        sol = CudaSolver(os.path.join(ROOT, 'generated_solvers', f'V_{name}.so'), spec)
        u, k, e, info = sol.solve_batch(small['x0'], small['xr'], small['ur'], arith=ARITH_EXACT, tail_mode=1)
        exact = bool(np.array_equal(u, ur_) and np.array_equal(k, kr) and np.array_equal(e, er))
        u, k, e, info = sol.solve_batch(small['x0'], small['xr'], small['ur'], arith=ARITH_EXACT, tail_mode=2, tail_grace=3)
        exact2 = bool(np.array_equal(u, ur_) and np.array_equal(k, kr) and np.array_equal(e, er))
        parked_small = info['parked']
        u, k, e, info = sol.solve_batch(small['x0'], small['xr'], small['ur'], arith=ARITH_FAST, tail_mode=2, tail_grace=3)
        same = k == kr
        relerr = float((np.abs(u[same] - ur_[same]) / np.maximum(1.0, np.abs(ur_[same]))).max())
        u1, k1, e1, _ = sol.solve_batch(small['x0'], small['xr'], small['ur'], arith=ARITH_FAST, tail_mode=1)
        fast = dict(e_same=bool(np.array_equal(e, er)), max_dk=int(np.abs(k - kr).max()), n_dk=int((~same).sum()), u_rel=relerr,
                    two_phase_same_bits=bool(np.array_equal(u.view(np.uint64), u1.view(np.uint64)) and np.array_equal(k, k1)),
                    parked=info['parked'])
        out[name] = dict(exact=exact, exact_two_phase=exact2, parked_small=parked_small, fast=fast, runs={})
        runs = [(1, 0, ()), (3, 0, (96, 320)), (3, 0, (128,))]
        if name == 'scalar':
            runs = [(1, 0, ()), (2, 32, ())]
        for mode, grace, caps in runs:
            ms = []
            for i in range(4):
                info = sol.solve_batch_device(B, d['x0'].data_ptr(), d['xr'].data_ptr(), d['ur'].data_ptr(), d_u.data_ptr(),
                                              d_k.data_ptr(), d_e.data_ptr(), tail_mode=mode, tail_grace=grace, tail_caps=caps)
                ms.append(info['kernel_ms'])
            out[name]['runs'][f'mode{mode}_g{grace}_c{"-".join(map(str, caps))}'] = dict(
                kernel_ms=min(ms[1:]), drain_us=info['drain_us'], span_us=info['span_us'], parked=info['parked'],
                launches=info['launches'], Msolves_s=B / min(ms[1:]) / 1e3)
        out[name].update(block=info['block_threads'], regs=info['regs_per_thread'], smem=info['smem_bytes'], sum_k=info['sum_k'])
        print(name, json.dumps(out[name]), flush=True)
        sol.free()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'variants_result.json'), 'w'), indent=1)


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20)
