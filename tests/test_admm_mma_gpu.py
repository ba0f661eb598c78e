"""GPU parity of the tensor-core (DMMA) engine of the equMPC ADMM solver (spcies_b200/csrc/MPC_ADMM_mma.cuh): 8 instances
per warp, v / lambda blocks in shared memory, the banded-Cholesky recurrences as merged FP64 MMA k-steps.
Gate (BASELINE.json north_star): e_flag identical, |dk| <= 1, u_opt <= 1e-9 relative on converged instances, against the
instantiated reference C solver (oracle/_ref)."""
import numpy as np
import pytest

from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, ENGINE_MMA, ENGINE_SCALAR, SpciesCudaError

pytestmark = pytest.mark.gpu
ADMM = ['T_equMPC_ADMM', 'C3_equMPC_ADMM', 'T_laxMPC_ADMM', 'T_ellipMPC_ADMM']   # equMPC, laxMPC (terminal block), ellipMPC (terminal ellipsoid)


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


@pytest.mark.parametrize('name', ADMM)
def test_admm_mma_engine_parity(name):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 6000, seed=51)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_MMA)
    _gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    u2, k2, e2, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_SCALAR)
    _gate(spec, u2, k2, e2, ur_, kr, er)
    # default engine = MMA: same bits as the explicit request
    u3, k3, e3, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'])
    assert np.array_equal(u3.view(np.uint64), u.view(np.uint64)) and np.array_equal(k3, k) and np.array_equal(e3, e)


def test_admm_mma_ragged_batches():
    sol, spec, cfg = prebuilt.get('C3_equMPC_ADMM')
    for B in (0, 1, 7, 8, 9, 63, 65, 257, 700):
        batch = sysmodel.synthetic_batch(cfg['sys'], max(B, 1), seed=52)
        x0, xr, ur = batch['x0'][:B], batch['xr'][:B], batch['ur'][:B]
        u, k, e, info = sol.solve_batch(x0, xr, ur, engine=ENGINE_MMA)
        assert u.shape == (B, sol.m)
        if B:
            ur_, kr, er = _ref('C3_equMPC_ADMM').solve_batch(x0, xr, ur, threads=8)
            _gate(spec, u, k, e, ur_, kr, er)


@pytest.mark.parametrize('name', ['T_equMPC_ADMM', 'T_laxMPC_ADMM'])
def test_admm_mma_per_instance_bounds(name):
    """opts.LB / UB: bounds equal to the generated constants reproduce the constant-bounds call bit for bit; random
    per-instance bounds agree with the EXACT scalar kernel (bit-identical to the reference for constant bounds) within the
    gate.  laxMPC: the terminal block takes the state part of the per-instance bounds."""
    _per_instance_bounds(name)


def _per_instance_bounds(name):
    sol, spec, cfg = prebuilt.get(name)
    B = 1200
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=53)
    LB = np.tile(np.concatenate([cfg['sys']['LBx'], cfg['sys']['LBu']]), (B, 1))
    UB = np.tile(np.concatenate([cfg['sys']['UBx'], cfg['sys']['UBu']]), (B, 1))
    r15 = np.vectorize(lambda v: float('%1.15f' % v))
    a = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA)
    b = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=r15(LB), UB=r15(UB), engine=ENGINE_MMA)
    assert np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64)) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    rng = np.random.default_rng(6)
    UB[:, :3] = rng.uniform(0.25, 0.35, size=(B, 3))
    LB[:, sol.n:] = -rng.uniform(0.5, 0.8, size=(B, sol.m))
    c = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=LB, UB=UB, engine=ENGINE_MMA)
    d = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=LB, UB=UB, arith=ARITH_EXACT)
    _gate(spec, c[0], c[1], c[2], d[0], d[1], d[2])


def test_admm_mma_engine_is_refused_where_it_cannot_run():
    """EXACT arithmetic and the debug payload belong to the scalar kernel: asking for MMA there is an error."""
    sol, spec, cfg = prebuilt.get('T_equMPC_ADMM')
    batch = sysmodel.synthetic_batch(cfg['sys'], 64, seed=54)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, engine=ENGINE_MMA)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, want_sol=True)
    sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT)
