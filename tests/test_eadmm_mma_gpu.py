"""GPU parity of the tensor-core (DMMA) engine of the MPCT EADMM solver (spcies_b200/csrc/MPCT_EADMM_mma.cuh): 8 instances
per warp, z1 / z3 / lambda / mu' blocks in shared memory, the banded-Cholesky recurrences as merged FP64 MMA k-steps.
Gate (BASELINE.json north_star): e_flag identical, |dk| <= 1, u_opt <= 1e-9 relative on converged instances, against the
instantiated reference C solver (oracle/_ref)."""
import numpy as np
import pytest

from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, ENGINE_MMA, ENGINE_SCALAR, SpciesCudaError

pytestmark = pytest.mark.gpu


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


from _parity import abs_err as _abs_err, rel_err as _rel_err   # true element-wise relative error (floor 1e-3), tests/_parity.py


def _gate(spec, u, k, e, ur_, kr, er):
    assert np.array_equal(e, er)
    assert np.max(np.abs(k - kr)) <= 1
    same = k == kr
    conv = er == 1
    assert _rel_err(u[same & conv], ur_[same & conv]) <= 1e-9
    assert _abs_err(u[same & ~conv], ur_[same & ~conv]) <= 1e-7      # instances that hit k_max: not a solution
    if (~same).any():
        assert _abs_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol'))


@pytest.mark.parametrize('name,B', [('T_MPCT_EADMM', 3000), ('C5b_MPCT_EADMM', 1500)])
def test_eadmm_mma_engine_parity(name, B):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=61)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_MMA)
    _gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    u2, k2, e2, _ = sol.solve_batch(batch['x0'][:512], batch['xr'][:512], batch['ur'][:512], arith=ARITH_FAST, engine=ENGINE_SCALAR)
    _gate(spec, u2, k2, e2, ur_[:512], kr[:512], er[:512])
    u3, k3, e3, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'])                    # default engine = MMA
    assert np.array_equal(u3.view(np.uint64), u.view(np.uint64)) and np.array_equal(k3, k) and np.array_equal(e3, e)


def test_eadmm_mma_ragged_batches():
    sol, spec, cfg = prebuilt.get('T_MPCT_EADMM')
    for B in (0, 1, 7, 8, 9, 63, 65, 257):
        batch = sysmodel.synthetic_batch(cfg['sys'], max(B, 1), seed=62)
        x0, xr, ur = batch['x0'][:B], batch['xr'][:B], batch['ur'][:B]
        u, k, e, info = sol.solve_batch(x0, xr, ur, engine=ENGINE_MMA)
        assert u.shape == (B, sol.m)
        if B:
            ur_, kr, er = _ref('T_MPCT_EADMM').solve_batch(x0, xr, ur, threads=8)
            _gate(spec, u, k, e, ur_, kr, er)


def test_eadmm_mma_engine_is_refused_where_it_cannot_run():
    sol, spec, cfg = prebuilt.get('T_MPCT_EADMM')
    batch = sysmodel.synthetic_batch(cfg['sys'], 64, seed=63)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, engine=ENGINE_MMA)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, want_sol=True)
    sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT)
