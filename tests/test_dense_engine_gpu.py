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
SOLVERS = {'T_MPCT_ADMM_cs': ('MPCT_ADMM', 'z'), 'T_MPCT_ADMM_semiband': None, 'T_HMPC_ADMM': None, 'T_ellipHMPC_ADMM': None}


def _inputs(sol, cfg, B, seed):
    """(x0, xr, ur) of a synthetic batch; solvers with three references (ellipHMPC) get small harmonic components."""
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=seed)
    if sol.nref == 1:
        return b['x0'], b['xr'], b['ur']
    rng = np.random.default_rng(seed + 7)
    n, m = sol.n, sol.m
    xr = (b['xr'], rng.uniform(-0.02, 0.02, (B, n)), rng.uniform(-0.02, 0.02, (B, n)))
    ur = (b['ur'], rng.uniform(-0.05, 0.05, (B, m)), rng.uniform(-0.05, 0.05, (B, m)))
    return b['x0'], xr, ur


def _cut(a, sl):
    return tuple(x[sl] for x in a) if isinstance(a, tuple) else a[sl]


def _status(sol, cfg):
    st = cfg['status']
    if sol.nref == 1:
        return st['x'], st['xr'], st['ur']
    return st['x'], (st['xr'], 0.01 * np.ones(sol.n), np.zeros(sol.n)), (st['ur'], np.zeros(sol.m), 0.02 * np.ones(sol.m))


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


@pytest.mark.parametrize('name', list(SOLVERS))
def test_single_instance_vs_golden(name, golden):
    sol, spec, cfg = prebuilt.get(name)
    x, xr, ur = _status(sol, cfg)
    u, k, e, s = sol.solve(x, xr, ur)
    ur_, kr, er, sr = _ref(name).solve(x, xr, ur)
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
    x0, xr, ur = _inputs(sol, cfg, 300, 31)
    u, k, e, info, s = sol.solve_batch(x0, xr, ur, arith=ARITH_EXACT, want_sol=True)
    ur_, kr, er, sr = _ref(name).solve_batch(x0, xr, ur, want_sol=True, threads=8)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    for f, _len in spec.sol_fields:
        assert np.array_equal(s[f], sr[f]), f
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())


@pytest.mark.parametrize('name', list(SOLVERS))
def test_dense_engine_parity(name):
    sol, spec, cfg = prebuilt.get(name)
    B = 3000
    x0, xr, ur = _inputs(sol, cfg, B, 32)
    ur_, kr, er = _ref(name).solve_batch(x0, xr, ur, threads=16)
    u, k, e, info = sol.solve_batch(x0, xr, ur, arith=ARITH_FAST, engine=ENGINE_MMA)
    gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    sl = slice(0, 256)
    u2, k2, e2, _ = sol.solve_batch(x0[sl], _cut(xr, sl), _cut(ur, sl), arith=ARITH_FAST, engine=ENGINE_SCALAR)
    gate(spec, u2, k2, e2, ur_[:256], kr[:256], er[:256])
    u3, k3, e3, _ = sol.solve_batch(x0, xr, ur)                                              # default engine = the dense engine
    assert np.array_equal(u3.view(np.uint64), u.view(np.uint64)) and np.array_equal(k3, k) and np.array_equal(e3, e)


@pytest.mark.parametrize('name', list(SOLVERS))
def test_dense_engine_ragged_batches(name):
    sol, spec, cfg = prebuilt.get(name)
    for B in (1, 7, 8, 9, 47, 49, 65, 130, 700):
        x0, xr, ur = _inputs(sol, cfg, B, 33 + B)
        u, k, e, info = sol.solve_batch(x0, xr, ur, engine=ENGINE_MMA)
        ur_, kr, er = _ref(name).solve_batch(x0, xr, ur, threads=8)
        gate(spec, u, k, e, ur_, kr, er)


@pytest.mark.parametrize('name', list(SOLVERS))
def test_dense_engine_refused_where_it_cannot_run(name):
    sol, spec, cfg = prebuilt.get(name)
    x0, xr, ur = _inputs(sol, cfg, 16, 34)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(x0, xr, ur, arith=ARITH_EXACT, engine=ENGINE_MMA)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(x0, xr, ur, engine=ENGINE_MMA, want_sol=True)


# ---- FISTA and ADMM solvers the banded engines do not take (other system sizes, penalty vectors, per-stage bounds): the generic
#      dense engine instead of the one-thread-per-instance kernel (MPC_FISTA_dense.cuh, MPC_ADMM_dense.cuh)
FALLBACK = ['S4_laxMPC_FISTA', 'S2_equMPC_ADMM', 'S4_equMPC_ADMM', 'S2_laxMPC_ADMM', 'S4_laxMPC_ADMM', 'T_equMPC_ADMM_vb',
            'T_equMPC_ADMM_vrho']


@pytest.mark.parametrize('name', FALLBACK)
def test_dense_engine_takes_what_the_banded_engines_do_not(name):
    sol, spec, cfg = prebuilt.get(name)
    B = 1500
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=66)
    ur_, kr, er = _ref(name).solve_batch(b['x0'], b['xr'], b['ur'], threads=16)
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], engine=ENGINE_MMA)            # an error if no tensor-core engine takes it
    gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    u2, k2, e2, info2 = sol.solve_batch(b['x0'], b['xr'], b['ur'])       # the default: this engine where it is the faster one (|z| >= 80)
    gate(spec, u2, k2, e2, ur_, kr, er)
    if name.startswith('S4_'):
        assert np.array_equal(u2.view(np.uint64), u.view(np.uint64)) and np.array_equal(k2, k) and info2['block_threads'] == info['block_threads']
    u3, k3, e3, info3 = sol.solve_batch(b['x0'][:256], b['xr'][:256], b['ur'][:256], engine=ENGINE_SCALAR)
    gate(spec, u3, k3, e3, ur_[:256], kr[:256], er[:256])
    ue, ke, ee, _ = sol.solve_batch(b['x0'][:256], b['xr'][:256], b['ur'][:256], arith=ARITH_EXACT)
    assert np.array_equal(ue.view(np.uint64), ur_[:256].view(np.uint64)) and np.array_equal(ke, kr[:256]) and np.array_equal(ee, er[:256])


@pytest.mark.parametrize('name', ['S4_laxMPC_FISTA', 'S4_equMPC_ADMM'])
def test_dense_fallback_ragged_batches(name):
    sol, spec, cfg = prebuilt.get(name)
    for B in (65, 127, 129, 700):
        b = sysmodel.synthetic_batch(cfg['sys'], B, seed=67)
        ur_, kr, er = _ref(name).solve_batch(b['x0'], b['xr'], b['ur'], threads=8)
        u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], engine=ENGINE_MMA)
        gate(spec, u, k, e, ur_, kr, er)
