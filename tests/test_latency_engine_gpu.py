"""GPU parity of the latency engine of the FISTA solvers (spcies_b200/csrc/MPC_FISTA_single.cuh): one CTA per instance, the
iteration of code_laxMPC_FISTA_C.c:323-389 restated in the space of the primal variable (one dense product with P = E' W^-1 E and
one barrier per iteration).  It serves the reference's own single-instance symbol (`laxMPC_FISTA(x0, xr, ur, u_opt, k, e_flag,
sol)`, header_laxMPC_FISTA_C.h:26; through a lingering one-CTA server kernel and a mailbox in mapped host memory) and host-buffer
batches of at most 64 instances (one launch, one CTA per instance).  Gate (BASELINE.json north_star): e_flag identical, |dk| <= 1, u_opt <= 1e-9 relative on converged
instances, against the instantiated reference C solver (oracle/_ref)."""
import numpy as np
import pytest

from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ENGINE_MMA, ENGINE_SCALAR, ENGINE_SINGLE, SpciesCudaError

from _parity import abs_err as _abs_err, rel_err as _rel_err

pytestmark = pytest.mark.gpu


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


def _gate(spec, u, k, e, ur_, kr, er):
    assert np.array_equal(e, er)
    assert np.max(np.abs(k - kr)) <= 1
    same = k == kr
    conv = er == 1
    assert _rel_err(u[same & conv], ur_[same & conv]) <= 1e-9
    assert _abs_err(u[same & ~conv], ur_[same & ~conv]) <= 1e-7
    if (~same).any():
        assert _abs_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol'))


@pytest.mark.parametrize('name', ['C2_laxMPC_FISTA', 'T_laxMPC_FISTA', 'T_equMPC_FISTA'])
def test_single_instance_symbol_matches_reference(name):
    """The unchanged single-instance symbol, one call per instance, against the reference's own function."""
    sol, spec, cfg = prebuilt.get(name)
    B = 96
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=91)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8)
    u = np.zeros_like(ur_)
    k = np.zeros(B, dtype=np.int64)
    e = np.zeros(B, dtype=np.int64)
    for i in range(B):
        ui, ki, ei, _ = sol.solve(batch['x0'][i], batch['xr'][i], batch['ur'][i])
        u[i], k[i], e[i] = ui, ki, ei
    _gate(spec, u, k, e, ur_, kr, er)


@pytest.mark.parametrize('name', ['C2_laxMPC_FISTA', 'T_equMPC_FISTA'])
def test_small_host_batches_take_the_latency_engine(name):
    sol, spec, cfg = prebuilt.get(name)
    batch = sysmodel.synthetic_batch(cfg['sys'], 80, seed=92)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8)
    for B in (1, 2, 7, 33, 64):
        u, k, e, info = sol.solve_batch(batch['x0'][:B], batch['xr'][:B], batch['ur'][:B])
        assert info['grid_blocks'] == B and info['block_threads'] % 32 == 0 and info['launches'] <= 1     # one CTA per instance (0 launches: the server was up)
        _gate(spec, u, k, e, ur_[:B], kr[:B], er[:B])
        assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
        u2, k2, e2, info2 = sol.solve_batch(batch['x0'][:B], batch['xr'][:B], batch['ur'][:B], engine=ENGINE_SINGLE)
        assert np.array_equal(u2.view(np.uint64), u.view(np.uint64)) and np.array_equal(k2, k) and np.array_equal(e2, e)
        u3, k3, e3, info3 = sol.solve_batch(batch['x0'][:B], batch['xr'][:B], batch['ur'][:B], engine=ENGINE_MMA)
        assert info3['grid_blocks'] != B or B == 1
        _gate(spec, u3, k3, e3, ur_[:B], kr[:B], er[:B])
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'])                 # 80 instances: a throughput engine
    assert info['grid_blocks'] < 80
    _gate(spec, u, k, e, ur_, kr, er)


def test_lingering_server_of_the_single_instance_symbol():
    """Single-instance calls are served by a one-CTA kernel that lingers 200 us after its last request (spcies_host.cuh:
    run_server): back-to-back calls reach the running kernel through its mailbox, a pause lets it exit and the next call starts
    a new one, a large batch in between tells it to go.  Every path returns the bits of the one-CTA-per-instance launch."""
    import time
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B = 40
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=95)
    u_l, k_l, e_l, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_SINGLE)       # launch path, one CTA each
    big = sysmodel.synthetic_batch(cfg['sys'], 4096, seed=96)
    for rnd in range(3):
        for i in range(B):
            u, k, e, _ = sol.solve(batch['x0'][i], batch['xr'][i], batch['ur'][i])
            assert np.array_equal(u.view(np.uint64), u_l[i].view(np.uint64)) and k == k_l[i] and e == e_l[i]
            if rnd == 1 and i % 8 == 0:
                time.sleep(0.002)                                                   # the server lingers out
            if rnd == 2 and i % 8 == 0:
                sol.solve_batch(big['x0'], big['xr'], big['ur'])                    # ... or is told to stop
    u1, k1, e1, info = sol.solve_batch(batch['x0'][:1], batch['xr'][:1], batch['ur'][:1])                 # a batch of one takes the same path
    assert np.array_equal(u1[0].view(np.uint64), u_l[0].view(np.uint64)) and info['grid_blocks'] == 1


def test_latency_engine_per_instance_bounds():
    from oracle import refs
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B = 60
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=93)
    variants = [refs.get_bounds_variant('C2_laxMPC_FISTA', s) for s in range(refs.N_BOUNDS_VARIANTS)]
    which = np.arange(B) % len(variants)
    LB = np.stack([np.concatenate([variants[w][1]['LBx'], variants[w][1]['LBu']]) for w in which])
    UB = np.stack([np.concatenate([variants[w][1]['UBx'], variants[w][1]['UBu']]) for w in which])
    r15 = np.vectorize(lambda v: float('%1.15f' % v))
    LB, UB = r15(LB), r15(UB)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=LB, UB=UB, engine=ENGINE_SINGLE)
    assert info['grid_blocks'] == B
    for s, (ref, _) in enumerate(variants):
        idx = np.nonzero(which == s)[0]
        ur_, kr, er = ref.solve_batch(batch['x0'][idx], batch['xr'][idx], batch['ur'][idx])
        _gate(spec, u[idx], k[idx], e[idx], ur_, kr, er)


def test_latency_engine_is_refused_where_it_cannot_run():
    sol, spec, cfg = prebuilt.get('T_laxMPC_FISTA')
    batch = sysmodel.synthetic_batch(cfg['sys'], 128, seed=94)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], engine=ENGINE_SINGLE)                      # more than 64 instances
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'][:8], batch['xr'][:8], batch['ur'][:8], arith=ARITH_EXACT, engine=ENGINE_SINGLE)
    with pytest.raises(SpciesCudaError):
        sol.solve_batch(batch['x0'][:8], batch['xr'][:8], batch['ur'][:8], engine=ENGINE_SINGLE, want_sol=True)
    # EXACT and the debug payload stay bit-identical to the reference on small batches (scalar kernel)
    ur_, kr, er = _ref('T_laxMPC_FISTA').solve_batch(batch['x0'][:8], batch['xr'][:8], batch['ur'][:8])
    u, k, e, _ = sol.solve_batch(batch['x0'][:8], batch['xr'][:8], batch['ur'][:8], arith=ARITH_EXACT)
    assert np.array_equal(u, ur_) and np.array_equal(k, kr) and np.array_equal(e, er)
    sol.solve_batch(batch['x0'][:8], batch['xr'][:8], batch['ur'][:8], engine=ENGINE_SCALAR)
