"""GPU parity of the FISTA kernels against the instantiated reference C solver (oracle/_ref)."""
import numpy as np
import pytest

from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST

pytestmark = pytest.mark.gpu

GOLD = {'T_laxMPC_FISTA': 'laxMPC_FISTA', 'T_equMPC_FISTA': 'equMPC_FISTA'}


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


def _rel_err(u, v):
    return np.max(np.abs(u - v) / np.maximum(1.0, np.abs(v)))


@pytest.mark.parametrize('name', ['T_laxMPC_FISTA', 'T_equMPC_FISTA'])
def test_single_instance_symbol_vs_golden_and_reference(name, golden):
    """Reference test point (tests/spcies_tester.m:114-116) through the UNCHANGED single-instance symbol."""
    sol, spec, cfg = prebuilt.get(name)
    st = cfg['status']
    u, k, e, s = sol.solve(st['x'], st['xr'], st['ur'])
    ur_, kr, er, sr = _ref(name).solve(st['x'], st['xr'], st['ur'])
    z_opt = np.array(golden[GOLD[name]]['z_opt'])
    assert e == er == 1
    assert abs(k - kr) <= 1
    assert np.max(np.abs(s['z'] - z_opt)) <= 1e-4            # tol_opt of tests/spcies_tester.m:261
    assert np.max(np.abs(s['z'] - sr['z'])) <= 1e-9
    assert np.max(np.abs(s['lambda'] - sr['lambda'])) <= 1e-9
    assert _rel_err(u, ur_) <= 1e-9


@pytest.mark.parametrize('name', ['T_laxMPC_FISTA', 'T_equMPC_FISTA', 'C2_laxMPC_FISTA'])
def test_exact_mode_is_bit_identical(name):
    """ARITH_EXACT: same IEEE operations in the same order as gcc -O3 => identical bits, k and e_flag."""
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 4096, seed=1)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8)
    assert np.array_equal(e, er)
    assert np.array_equal(k, kr)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    assert info['sum_k'] == int(k.sum())
    assert info['n_not_converged'] == int((e == -1).sum())


@pytest.mark.parametrize('name', ['T_laxMPC_FISTA', 'T_equMPC_FISTA', 'C2_laxMPC_FISTA'])
def test_fast_mode_parity(name):
    """ARITH_FAST (FMA): u_opt within 1e-9 relative, e_flag identical, |dk| <= 1 (BASELINE.json north_star)."""
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 8192, seed=2)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8)
    assert np.array_equal(e, er)
    assert np.max(np.abs(k - kr)) <= 1
    same = k == kr
    assert _rel_err(u[same], ur_[same]) <= 1e-9
    # an instance whose k moved by one stops one iterate earlier/later: compare at the solver tolerance
    if (~same).any():
        assert _rel_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol'))


def test_debug_payload_batch():
    sol, spec, cfg = prebuilt.get('T_laxMPC_FISTA')
    batch = sysmodel.synthetic_batch(cfg['sys'], 257, seed=3)
    u, k, e, info, s = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, want_sol=True)
    ur_, kr, er, sr = _ref('T_laxMPC_FISTA').solve_batch(batch['x0'], batch['xr'], batch['ur'], want_sol=True)
    assert np.array_equal(s['z'], sr['z'])
    assert np.array_equal(s['lambda'], sr['lambda'])


def test_per_instance_bounds_match_regenerated_reference():
    """Per-instance bounds (opts.LB/UB): every distinct bound set must reproduce a reference solver that was
    generated with those bounds as constants."""
    from oracle import refs
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B = 96
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=4)
    n, m = sol.n, sol.m
    variants = [refs.get_bounds_variant('C2_laxMPC_FISTA', s) for s in range(refs.N_BOUNDS_VARIANTS)]
    which = np.arange(B) % len(variants)
    LB = np.stack([np.concatenate([variants[w][1]['LBx'], variants[w][1]['LBu']]) for w in which])
    UB = np.stack([np.concatenate([variants[w][1]['UBx'], variants[w][1]['UBu']]) for w in which])
    # the regenerated reference sees its bounds through the generator's %1.15f text (dec_var.m:259)
    r15 = np.vectorize(lambda v: float('%1.15f' % v))
    LB, UB = r15(LB), r15(UB)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=LB, UB=UB, arith=ARITH_EXACT)
    for s, (ref, _) in enumerate(variants):
        idx = np.nonzero(which == s)[0]
        ur_, kr, er = ref.solve_batch(batch['x0'][idx], batch['xr'][idx], batch['ur'][idx])
        assert np.array_equal(k[idx], kr) and np.array_equal(e[idx], er)
        assert np.array_equal(u[idx], ur_)


def test_empty_and_ragged_batches():
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    for B in (0, 1, 31, 33, 127, 129, 1000):
        batch = sysmodel.synthetic_batch(cfg['sys'], max(B, 1), seed=6)
        x0, xr, ur = batch['x0'][:B], batch['xr'][:B], batch['ur'][:B]
        u, k, e, info = sol.solve_batch(x0, xr, ur, arith=ARITH_EXACT)
        assert u.shape == (B, sol.m)
        if B:
            ur_, kr, er = _ref('C2_laxMPC_FISTA').solve_batch(x0, xr, ur)
            assert np.array_equal(u, ur_) and np.array_equal(k, kr) and np.array_equal(e, er)
