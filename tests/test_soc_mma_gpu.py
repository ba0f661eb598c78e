"""GPU parity of the tensor-core (DMMA) engines of the ellipMPC ADMM_soc solver: the structured one
(spcies_b200/csrc/ellipMPC_ADMM_soc_band.cuh: W is block tridiagonal with N + 1 blocks once the two rows that pin t are taken
out, so the iteration is the equMPC ADMM engine with a dense terminal block and a cone block) and the dense one
(spcies_b200/csrc/ellipMPC_ADMM_soc_mma.cuh: the reference's CSR / CSC-LDL chain folded into one dense linear map applied as a
batched FP64 MMA GEMM; kept for the shapes / structures the first does not take, SPCIES_CUDA_SOC_ENGINE=dense forces it).
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
        assert _abs_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol_p'))


@pytest.mark.parametrize('name,B', [('T_ellipMPC_ADMM_soc', 2000), ('C4_ellipMPC_ADMM_soc', 6000)])
def test_soc_mma_engine_parity(name, B):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=81, with_r=True)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], arith=ARITH_FAST, engine=ENGINE_MMA)
    _gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    nb = 512
    u2, k2, e2, _ = sol.solve_batch(batch['x0'][:nb], batch['xr'][:nb], batch['ur'][:nb], r=batch['r'][:nb], arith=ARITH_FAST,
                                    engine=ENGINE_SCALAR)
    _gate(spec, u2, k2, e2, ur_[:nb], kr[:nb], er[:nb])
    u3, k3, e3, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'])       # default engine = MMA
    assert np.array_equal(u3.view(np.uint64), u.view(np.uint64)) and np.array_equal(k3, k) and np.array_equal(e3, e)


@pytest.mark.parametrize('name', ['T_ellipMPC_ADMM_soc', 'C4_ellipMPC_ADMM_soc'])
def test_soc_structured_and_dense_engines_agree(name, monkeypatch):
    """Both engines against the reference on the same batch; the default is the structured one (512 threads per CTA; the
    dense one sizes its CTA by the shared memory its fragment table leaves)."""
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 3000, seed=84, with_r=True)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], engine=ENGINE_MMA)
    _gate(spec, u, k, e, ur_, kr, er)
    monkeypatch.setenv('SPCIES_CUDA_SOC_ENGINE', 'dense')
    ud, kd, ed, infod = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], engine=ENGINE_MMA)
    monkeypatch.delenv('SPCIES_CUDA_SOC_ENGINE')
    _gate(spec, ud, kd, ed, ur_, kr, er)
    assert info['block_threads'] == 512 and infod['block_threads'] != 512
    assert np.array_equal(e, ed) and np.max(np.abs(k - kd)) <= 1


def test_soc_mma_ragged_batches():
    sol, spec, cfg = prebuilt.get('C4_ellipMPC_ADMM_soc')
    for B in (0, 1, 7, 8, 9, 55, 57, 225, 500):
        batch = sysmodel.synthetic_batch(cfg['sys'], max(B, 1), seed=82, with_r=True)
        x0, xr, ur, r = batch['x0'][:B], batch['xr'][:B], batch['ur'][:B], batch['r'][:B]
        u, k, e, info = sol.solve_batch(x0, xr, ur, r=r, engine=ENGINE_MMA) if B else sol.solve_batch(x0, xr, ur, r=np.zeros(1))
        assert u.shape == (B, sol.m)
        if B:
            ur_, kr, er = _ref('C4_ellipMPC_ADMM_soc').solve_batch(x0, xr, ur, r=r, threads=8)
            _gate(spec, u, k, e, ur_, kr, er)


def test_soc_mma_engine_is_refused_where_it_cannot_run():
    sol, spec, cfg = prebuilt.get('T_ellipMPC_ADMM_soc')
    batch = sysmodel.synthetic_batch(cfg['sys'], 32, seed=83, with_r=True)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], arith=ARITH_EXACT, engine=ENGINE_MMA)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], engine=ENGINE_MMA, want_sol=True)
    sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], r=batch['r'], arith=ARITH_EXACT)
