"""GPU parity of the tensor-core (DMMA) engine of the HMPC ADMM_split / SADMM_split solvers
(spcies_b200/csrc/HMPC_ADMM_split_mma.cuh): the dense product primal_hat = M2 bh - M1 q_hat as a batched GEMM (8 instances
per warp, [-M1 | M2] as FP64 MMA B fragments streamed from L2), diamond-set projections on lane-local triples.
Gate (BASELINE.json north_star): e_flag identical, |dk| <= 1, u_opt <= 1e-9 relative on converged instances, against the
instantiated reference C solver (oracle/_ref; HMPC parity is pinned by the instantiated template only)."""
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
        assert _abs_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol_p'))


@pytest.mark.parametrize('name,B', [('T_HMPC_ADMM_split', 1500), ('T_HMPC_SADMM_split', 1500), ('T_HMPC_SADMM_split_sparse', 600),
                                    ('C5a_HMPC_SADMM_split', 160)])
def test_hmpc_mma_engine_parity(name, B):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=71)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_MMA)
    _gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    nb = min(B, 256)
    u2, k2, e2, _ = sol.solve_batch(batch['x0'][:nb], batch['xr'][:nb], batch['ur'][:nb], arith=ARITH_FAST, engine=ENGINE_SCALAR)
    _gate(spec, u2, k2, e2, ur_[:nb], kr[:nb], er[:nb])
    u3, k3, e3, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'])                    # default engine = MMA
    assert np.array_equal(u3.view(np.uint64), u.view(np.uint64)) and np.array_equal(k3, k) and np.array_equal(e3, e)


def test_hmpc_mma_ragged_batches():
    sol, spec, cfg = prebuilt.get('T_HMPC_SADMM_split')
    for B in (0, 1, 7, 8, 9, 63, 65, 130):
        batch = sysmodel.synthetic_batch(cfg['sys'], max(B, 1), seed=72)
        x0, xr, ur = batch['x0'][:B], batch['xr'][:B], batch['ur'][:B]
        u, k, e, info = sol.solve_batch(x0, xr, ur, engine=ENGINE_MMA)
        assert u.shape == (B, sol.m)
        if B:
            ur_, kr, er = _ref('T_HMPC_SADMM_split').solve_batch(x0, xr, ur, threads=8)
            _gate(spec, u, k, e, ur_, kr, er)


def test_hmpc_mma_engine_is_refused_where_it_cannot_run():
    sol, spec, cfg = prebuilt.get('T_HMPC_ADMM_split')
    batch = sysmodel.synthetic_batch(cfg['sys'], 32, seed=73)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, engine=ENGINE_MMA)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, want_sol=True)
    sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT)
