"""GPU: the batched, device-resident closed loop (`<func>_closed_loop`) against a loop of reference calls
(examples/cl_in_C/main_cl_in_C.c:100-117 for every instance, oracle/ref_batch_driver.c) and -- for the warm start -- against the
dense NumPy twin of the reference's MATLAB solver with its `lambda` argument (tests/fista_twin.py)."""
import numpy as np
import pytest

import fista_twin
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, ENGINE_SCALAR, SpciesCudaError

pytestmark = pytest.mark.gpu


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


def _AB(cfg):
    return np.hstack([cfg['sys']['A'], cfg['sys']['B']])


@pytest.mark.parametrize('name', ['C2_laxMPC_FISTA', 'T_equMPC_ADMM', 'T_MPCT_EADMM', 'C4_ellipMPC_ADMM_soc'])
def test_closed_loop_exact_is_a_loop_of_reference_calls(name):
    """EXACT arithmetic, cold start: states, inputs, iteration counts and flags of every sampling time are bit-identical to the
    loop of reference calls (solver launches on device-resident arrays, plant step accumulated in the example's order)."""
    sol, spec, cfg = prebuilt.get(name)
    B, steps = 96, 6
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=71, with_r=sol.has_r)
    kw = dict(r=b['r']) if sol.has_r else {}
    xr_, ur_, kr, er = _ref(name).closed_loop(b['x0'], b['xr'], b['ur'], steps, _AB(cfg), threads=8, **kw)
    x, u, k, e, info = sol.closed_loop(b['x0'], b['xr'], b['ur'], steps, arith=ARITH_EXACT, **kw)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    assert np.array_equal(x.view(np.uint64), xr_.view(np.uint64))
    assert info['sum_k'] == int(k.sum())


def test_closed_loop_on_chip_cold_start():
    """FAST arithmetic on the FISTA tensor-core engine: the whole run is ONE launch, every instance stays on chip across the
    sampling times (successor state = one more MMA).  Against the loop of reference calls: e_flag identical, |dk| <= 1, and the
    trajectories within 1e-9 where the iteration counts agree all along (a step that stops one iterate earlier / later moves the
    rest of that instance's trajectory by the solver tolerance)."""
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B, steps = 4000, 12
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=72)
    xr_, ur_, kr, er = _ref('C2_laxMPC_FISTA').closed_loop(b['x0'], b['xr'], b['ur'], steps, _AB(cfg), threads=16)
    x, u, k, e, info = sol.closed_loop(b['x0'], b['xr'], b['ur'], steps)
    assert info['launches'] == 1
    assert info['sum_k'] == int(k.sum())
    same = np.all(k == kr, axis=0)                      # instances whose iteration counts agree at every sampling time
    assert same.mean() > 0.97
    assert np.array_equal(e[:, same], er[:, same])
    assert np.max(np.abs(u[:, same] - ur_[:, same])) <= 1e-9
    assert np.max(np.abs(x[:, same] - xr_[:, same])) <= 1e-9
    assert np.max(np.abs(u - ur_)) <= 20 * float(spec.define('tol'))
    # the launch-per-sampling-time path (scalar engine) gives the same answer
    x2, u2, k2, e2, info2 = sol.closed_loop(b['x0'][:512], b['xr'][:512], b['ur'][:512], steps, engine=ENGINE_SCALAR)
    assert info2['launches'] == 2 * steps
    s2 = np.all(k2 == kr[:, :512], axis=0)
    assert np.max(np.abs(u2[:, s2] - ur_[:, :512][:, s2])) <= 1e-9


@pytest.mark.parametrize('warm', [1, 2])
def test_closed_loop_warm_start_matches_the_matlab_twin_semantics(warm):
    """warm_start = 1: every sampling time starts from the dual point of the previous exit test -- the `lambda` argument of
    platforms/Matlab/spcies_laxMPC_FISTA_solver.m:161-164 fed with the previous sol.lambda; warm_start = 2: that point shifted by
    one stage (the horizon receded).  Checked against the dense NumPy twin run the same way.  (On this problem the unshifted
    start costs 6 % more iterations than the cold start, the shifted one 9 % less -- measured with the twin.)"""
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B, steps = 48, 10
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=73)
    P = fista_twin.build(cfg['sys'], cfg['param'])
    tol, kmax = float(spec.define('tol')), int(spec.define('k_max'))
    xw, uw, kw_, ew, iw = sol.closed_loop(b['x0'], b['xr'], b['ur'], steps, warm_start=warm)
    xc, uc, kc, ec, ic = sol.closed_loop(b['x0'], b['xr'], b['ur'], steps, warm_start=0)
    n_same, k_twin = 0, 0
    for i in range(B):
        xt, ut, kt, et = fista_twin.closed_loop(P, b['x0'][i], b['xr'][i], b['ur'][i], steps, warm, tol, kmax)
        k_twin += int(kt.sum())
        if np.array_equal(kt, kw_[:, i]):
            n_same += 1
            assert np.array_equal(et, ew[:, i])
            assert np.max(np.abs(ut - uw[:, i])) <= 1e-8 and np.max(np.abs(xt - xw[:, i])) <= 1e-8
        else:
            assert np.max(np.abs(ut - uw[:, i])) <= 50 * tol
    assert n_same >= 0.8 * B
    assert abs(iw['sum_k'] - k_twin) <= 0.02 * k_twin
    if warm == 2:
        assert iw['sum_k'] < ic['sum_k']                # the shifted warm start saves iterations
    with pytest.raises(SpciesCudaError):
        sol.closed_loop(b['x0'], b['xr'], b['ur'], steps, warm_start=warm, arith=ARITH_EXACT)


def test_closed_loop_plant_mismatch_and_sharding_arguments():
    """A plant other than the prediction model (opts.plant_AB) uses the launch-per-sampling-time path; the result is the loop of
    reference calls with that plant."""
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B, steps = 64, 5
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=74)
    AB = _AB(cfg) * (1.0 + 0.01 * np.random.default_rng(3).standard_normal((sol.n, sol.n + sol.m)))
    xr_, ur_, kr, er = _ref('C2_laxMPC_FISTA').closed_loop(b['x0'], b['xr'], b['ur'], steps, AB, threads=8)
    x, u, k, e, info = sol.closed_loop(b['x0'], b['xr'], b['ur'], steps, plant_AB=AB, arith=ARITH_EXACT)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64)) and np.array_equal(x.view(np.uint64), xr_.view(np.uint64))
