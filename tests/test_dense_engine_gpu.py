"""GPU parity of the solvers added in round 2 on the generic dense tensor-core engine (spcies_b200/csrc/spcies_dense_mma.cuh:
the solver's sparse chain folded into one linear map, streamed L2 -> shared memory by a producer warp with bulk async copies and
applied as a batched FP64 MMA GEMM) and of their bit-exact one-thread-per-instance kernels, against the instantiated reference
C templates (oracle/_ref)."""
import numpy as np
import pytest

from _parity import gate
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, ENGINE_MMA, ENGINE_SCALAR, SpciesCudaError

pytestmark = pytest.mark.gpu

# solver -> (golden vector of the reference's own test, sol field) or None
SOLVERS = {'T_MPCT_ADMM_cs': ('MPCT_ADMM', 'z')}


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


@pytest.mark.parametrize('name', list(SOLVERS))
def test_single_instance_vs_golden(name, golden):
    sol, spec, cfg = prebuilt.get(name)
    st = cfg['status']
    u, k, e, s = sol.solve(st['x'], st['xr'], st['ur'])
    ur_, kr, er, sr = _ref(name).solve(st['x'], st['xr'], st['ur'])
    assert e == er == 1 and abs(k - kr) <= 1
    if SOLVERS[name] is not None:
        gold_name, field = SOLVERS[name]
        z_opt = np.array(golden[gold_name]['z_opt'])
        assert np.max(np.abs(s[field][:len(z_opt)] - z_opt)) <= 1e-4        # tol_opt of tests/spcies_tester.m:261
    for f, _len in spec.sol_fields:
        assert np.max(np.abs(s[f] - sr[f])) <= 1e-9, f
    assert np.max(np.abs(u - ur_)) <= 1e-9


@pytest.mark.parametrize('name', list(SOLVERS))
def test_exact_mode_bit_identical_with_debug_payload(name):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 300, seed=31)
    u, k, e, info, s = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, want_sol=True)
    ur_, kr, er, sr = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], want_sol=True, threads=8)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    for f, _len in spec.sol_fields:
        assert np.array_equal(s[f], sr[f]), f
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())


@pytest.mark.parametrize('name', list(SOLVERS))
def test_dense_engine_parity(name):
    sol, spec, cfg = prebuilt.get(name)
    B = 3000
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=32)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_MMA)
    gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    u2, k2, e2, _ = sol.solve_batch(batch['x0'][:256], batch['xr'][:256], batch['ur'][:256], arith=ARITH_FAST, engine=ENGINE_SCALAR)
    gate(spec, u2, k2, e2, ur_[:256], kr[:256], er[:256])
    u3, k3, e3, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'])                   # default engine = the dense engine
    assert np.array_equal(u3.view(np.uint64), u.view(np.uint64)) and np.array_equal(k3, k) and np.array_equal(e3, e)


@pytest.mark.parametrize('name', list(SOLVERS))
def test_dense_engine_ragged_batches(name):
    sol, spec, cfg = prebuilt.get(name)
    for B in (1, 7, 8, 9, 47, 49, 65, 130, 700):
        batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=33 + B)
        u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA)
        ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8)
        gate(spec, u, k, e, ur_, kr, er)


@pytest.mark.parametrize('name', list(SOLVERS))
def test_dense_engine_refused_where_it_cannot_run(name):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 16, seed=34)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, engine=ENGINE_MMA)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, want_sol=True)
