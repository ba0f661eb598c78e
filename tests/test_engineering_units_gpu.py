"""GPU: options.in_engineering -- inputs and outputs in engineering units.  The generated solver scales (x0, xr, ur) with the
system's scaling vectors and operating point on the way in and u_opt on the way out (code_laxMPC_FISTA_C.c:75-82, :398-402; the
same block in every template), in the template's operation order: EXACT is bit-identical to the reference built with
`#define in_engineering 1`, the tensor-core engines (which scale at refill and at write-back) meet the FAST gate.  One solver per
kernel family: banded FISTA (+ latency engine / server), banded ADMM, structured ADMM_soc, EADMM, the dense engine."""
import numpy as np
import pytest

from _parity import gate
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, ENGINE_MMA

pytestmark = pytest.mark.gpu

ENG = [k for k in prebuilt.SOLVERS if k.startswith('E_')]


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


def _engineering_batch(sol, cfg, B, seed):
    """A synthetic batch in scaled units, expressed in engineering units: x_in = x / Nx + x_op."""
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=seed, with_r=sol.has_r)
    s = cfg['sys']
    Nx, Nu, xo, uo = (np.asarray(s[k], float) for k in ('Nx', 'Nu', 'x0', 'u0'))
    out = dict(x0=b['x0'] / Nx + xo, xr=b['xr'] / Nx + xo, ur=b['ur'] / Nu + uo)
    if sol.has_r:
        out['r'] = b['r']
    return out


@pytest.mark.parametrize('name', ENG)
def test_engineering_units_exact_and_fast(name):
    sol, spec, cfg = prebuilt.get(name)
    assert int(spec.define('in_engineering')) == 1
    B = 600
    b = _engineering_batch(sol, cfg, B, seed=71)
    kw = dict(r=b['r']) if sol.has_r else {}
    ur_, kr, er = _ref(name).solve_batch(b['x0'], b['xr'], b['ur'], threads=16, **kw)
    assert (er == 1).mean() > 0.5 and np.abs(ur_ - np.asarray(cfg['sys']['u0'])).max() > 0.05       # a meaningful batch
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], arith=ARITH_EXACT, **kw)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], arith=ARITH_FAST, engine=ENGINE_MMA, **kw)      # the tensor-core engine
    gate(spec, u, k, e, ur_, kr, er)
    u, k, e, info = sol.solve_batch(b['x0'][:40], b['xr'][:40], b['ur'][:40], **({'r': b['r'][:40]} if sol.has_r else {}))   # small-call path
    gate(spec, u, k, e, ur_[:40], kr[:40], er[:40])
    for i in range(4):                                                                              # the single-instance symbol
        us, ks, es, _ = sol.solve(b['x0'][i], b['xr'][i], b['ur'][i], **({'r': b['r'][i]} if sol.has_r else {}))
        assert es == er[i] and abs(ks - kr[i]) <= 1
        if ks == kr[i] and es == 1:
            assert np.max(np.abs(us - ur_[i])) <= 1e-9 * max(1.0, np.abs(ur_[i]).max())


def test_engineering_units_closed_loop():
    """The closed loop of an in_engineering solver simulates a plant given in engineering units (the prediction model acts on
    scaled, incremental variables, so it is not a default): without one the call is refused, with one it is the loop of
    reference calls around the same plant."""
    from spcies_b200.solver import SpciesCudaError
    name = 'E_laxMPC_FISTA'
    sol, spec, cfg = prebuilt.get(name)
    B, steps = 64, 6
    b = _engineering_batch(sol, cfg, B, seed=72)
    with pytest.raises(SpciesCudaError):
        sol.closed_loop(b['x0'], b['xr'], b['ur'], steps)
    AB = np.hstack([cfg['sys']['A'], cfg['sys']['B']]) * 0.98           # some plant, in engineering units
    x, u, k, e, info = sol.closed_loop(b['x0'], b['xr'], b['ur'], steps, plant_AB=AB, arith=ARITH_EXACT)
    from oracle import refs
    xr_, ur_, kr, er = refs.get(name)[0].closed_loop(b['x0'], b['xr'], b['ur'], steps, AB)
    assert np.array_equal(e, er) and np.array_equal(k, kr)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64)) and np.array_equal(x.view(np.uint64), xr_.view(np.uint64))
