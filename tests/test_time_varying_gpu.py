"""GPU: options.time_varying -- every instance brings its own model A, B, Q, R and bounds, and the block-Cholesky
factorisation of its W runs on the device (csrc/MPC_FISTA_tv.cuh, code_laxMPC_FISTA_C.c:101-272; csrc/MPC_ADMM_tv.cuh,
code_equMPC_ADMM_C.c:100-264).  Oracle: the reference's own
template instantiated with `#define TIME_VARYING 1` and called with the same per-instance arguments."""
import numpy as np
import pytest

from _parity import gate
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST

pytestmark = pytest.mark.gpu


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


TV = {'TV_laxMPC_FISTA': 'C2_laxMPC_FISTA', 'TV_equMPC_ADMM': 'C3_equMPC_ADMM'}     # time-varying solver -> its constant-model twin


@pytest.mark.parametrize('name', list(TV) + ['TV_equMPC_FISTA', 'TV_laxMPC_ADMM'])
def test_time_varying_batch_exact_and_fast(name):
    sol, spec, cfg = prebuilt.get(name)
    assert sol.tv
    B = 700
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=51)
    tv, LB, UB = sysmodel.perturbed_models(cfg['sys'], cfg['param'], B, seed=52)
    ur_, kr, er, sr = _ref(name).solve_batch(b['x0'], b['xr'], b['ur'], tv=tv, LB=LB, UB=UB, threads=16, want_sol=True)
    assert len(np.unique(kr)) > 10                       # the models really differ
    u, k, e, info, s = sol.solve_batch(b['x0'], b['xr'], b['ur'], tv=tv, LB=LB, UB=UB, arith=ARITH_EXACT, want_sol=True)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    for f, _len in spec.sol_fields:
        assert np.array_equal(s[f], sr[f]), f
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], tv=tv, LB=LB, UB=UB, arith=ARITH_FAST)
    gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum())


@pytest.mark.parametrize('name', list(TV))
def test_time_varying_single_instance_symbol_and_nominal_model(name):
    """The reference signature (A_in, B_in column-major, Q_in, R_in, LB_in, UB_in) through the single-instance symbol; with the
    nominal model the answer is the constant-model solver's (the device factorisation reproduces the generator's Alpha / Beta)."""
    sol, spec, cfg = prebuilt.get(name)
    s, st = cfg['sys'], cfg['status']
    tv = (s['A'][None], s['B'][None], np.diag(cfg['param']['Q'])[None], np.diag(cfg['param']['R'])[None])
    LB, UB = np.concatenate([s['LBx'], s['LBu']])[None], np.concatenate([s['UBx'], s['UBu']])[None]
    u, k, e, so = sol.solve(st['x'], st['xr'], st['ur'], tv=tv, LB=LB, UB=UB)
    c2, _, _ = prebuilt.get(TV[name])
    u0, k0, e0, _ = c2.solve(st['x'], st['xr'], st['ur'])
    assert e == e0 == 1 and k == k0
    assert np.max(np.abs(u - u0)) <= 1e-9
