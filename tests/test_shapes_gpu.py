"""GPU parity on systems of other dimensions than the 3-mass benchmark system (+sp_utils/gen_oscillating_masses.m:28-59 with
2 masses: n = 4, m = 2; 4 masses, all actuated: n = 8, m = 4 = more than one 8-column MMA tile per stage vector; and the odd
sizes n = 5, m = 2 / n = 7, m = 1 obtained by dropping a velocity state).  The FP64
tensor-core engines take 5 <= n <= 6, n + m <= 8 (the FISTA engine: any n + m <= 8 with its plain column layout; the dense
engines: any shape); everything else must fall back to the one-thread-per-instance kernels *and still match the reference*:
a generator that silently emits a wrong or unbuildable solver for another system is what these tests catch."""
import numpy as np
import pytest

from _parity import gate
from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, ENGINE_MMA, SpciesCudaError

pytestmark = pytest.mark.gpu

SHAPE_SOLVERS = [k for k in prebuilt.SOLVERS if k[:3] in ('S2_', 'S4_', 'S5_', 'S7_')]
# the odd sizes: banded engines at n = 5 (MERGE layout), the plain FISTA layout at n = 7, m = 1; n = 7 ADMM: the dense engine; n = 7
# EADMM has none (its engine takes 5 <= n <= 6)
HAS_ENGINE = {'S5_laxMPC_FISTA', 'S7_laxMPC_FISTA', 'S5_equMPC_ADMM', 'S7_equMPC_ADMM', 'S5_MPCT_EADMM'}


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


@pytest.mark.parametrize('name', SHAPE_SOLVERS)
def test_other_dimensions_exact_and_fast(name):
    sol, spec, cfg = prebuilt.get(name)
    assert (sol.n, sol.m) in ((4, 2), (8, 4), (5, 2), (7, 1))
    B = 400
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=61, with_r=sol.has_r)
    kw = dict(r=b['r']) if sol.has_r else {}
    ur_, kr, er = _ref(name).solve_batch(b['x0'], b['xr'], b['ur'], threads=8, **kw)
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], arith=ARITH_EXACT, **kw)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], arith=ARITH_FAST, **kw)      # whatever engine AUTO picks for this shape
    gate(spec, u, k, e, ur_, kr, er)
    assert info['sum_k'] == int(k.sum())
    try:                                                                                     # the tensor-core engine, where one takes this shape
        u, k, e, info = sol.solve_batch(b['x0'], b['xr'], b['ur'], arith=ARITH_FAST, engine=ENGINE_MMA, **kw)
    except SpciesCudaError:
        assert name not in HAS_ENGINE, name
    else:
        gate(spec, u, k, e, ur_, kr, er)
        assert name in HAS_ENGINE or not name.startswith(('S5_', 'S7_')), name
