#!/usr/bin/env python
"""Development tool: build one prebuilt solver under several sets of compile-time switches and time each on one B200.

    python tools/engine_variants.py build C4_ellipMPC_ADMM_soc b384=-DSPCIES_SOC_BAND_BLOCK=384 b512=-DSPCIES_SOC_BAND_BLOCK=512
    python tools/engine_variants.py run   C4_ellipMPC_ADMM_soc [B]       # on the GPU box

`build` (here, nvcc cross-compiles) writes generated_solvers/V_<config>_<variant>.so and the list of variants; `run` times the
default FAST path of each on B instances resident on the device (best of 3) and checks k / e_flag against the first variant.
Several flags of one variant are separated by commas."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spcies_b200 import prebuilt, sysmodel                      # noqa: E402
from spcies_b200.platforms import cuda_code                     # noqa: E402


def _list(cfg_name):
    return os.path.join(ROOT, 'generated_solvers', f'V_{cfg_name}.variants.json')


def build(cfg_name, variants):
    spec, cfg = prebuilt.spec_for(cfg_name)
    done = {'default': []}
    done.update({k: v.split(',') for k, v in (a.split('=', 1) for a in variants)})
    for name, flags in done.items():
        cu, _ = cuda_code.emit(spec, save_name=f'V_{cfg_name}_{name}')
        cuda_code.build(cu, extra_flags=tuple(flags))
        log = open(cu[:-3] + '.stamp').read()
        print(name, flags, [l.strip() for l in log.split('\n') if 'Used' in l or 'spill' in l][-2:], flush=True)
    json.dump(done, open(_list(cfg_name), 'w'))


def run(cfg_name, B):
    import numpy as np
    import torch
    from spcies_b200.solver import CudaSolver
    spec, cfg = prebuilt.spec_for(cfg_name)
    with_r = 'r_ellip' in spec.extra_inputs
    big = sysmodel.synthetic_batch(cfg['sys'], B, seed=100, **({'with_r': True} if with_r else {}))
    dev = torch.device('cuda', 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in big.items()}
    d_u = torch.empty((B, spec.dims['m']), dtype=torch.float64, device=dev)
    d_k = torch.empty(B, dtype=torch.int32, device=dev)
    d_e = torch.empty(B, dtype=torch.int32, device=dev)
    first = None
    for name in json.load(open(_list(cfg_name))):
        sol = CudaSolver(os.path.join(ROOT, 'generated_solvers', f'V_{cfg_name}_{name}.so'), spec)
        kw = {'d_r': d['r'].data_ptr()} if with_r else {}
        ms = []
        for _ in range(4):
            info = sol.solve_batch_device(B, d['x0'].data_ptr(), d['xr'].data_ptr(), d['ur'].data_ptr(), d_u.data_ptr(), d_k.data_ptr(),
                                          d_e.data_ptr(), **kw)
            ms.append(info['kernel_ms'])
        k, e = d_k.cpu().numpy(), d_e.cpu().numpy()
        first = first or (k, e)
        print(name, 'kernel_ms %.3f' % min(ms[1:]), 'Msolves/s %.3f' % (B / min(ms[1:]) / 1e3), 'block', info['block_threads'],
              'regs', info['regs_per_thread'], 'smem', info['smem_bytes'], 'k,e same as first:',
              bool((k == first[0]).all() and (e == first[1]).all()), flush=True)
        sol.free()


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build(sys.argv[2], sys.argv[3:])
    else:
        run(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20)
