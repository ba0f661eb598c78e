"""CPU tier: the oracle (instantiated reference templates, oracle/_ref) against every golden vector and
fixture the reference's own tests hold, plus the host-side model restatement."""
import numpy as np
import pytest

from spcies_b200 import prebuilt, sysmodel

GOLD = {'T_laxMPC_FISTA': ('laxMPC_FISTA', 'z', 54), 'T_equMPC_FISTA': ('equMPC_FISTA', 'z', 66),
        'T_laxMPC_ADMM': ('laxMPC_ADMM', 'z', 1264), 'T_equMPC_ADMM': ('equMPC_ADMM', 'z', 1271),
        'T_ellipMPC_ADMM': ('ellipMPC_ADMM', 'z', 1278), 'T_ellipMPC_ADMM_soc': ('ellipMPC_ADMM_soc', 'z', 879),
        'T_MPCT_EADMM': ('MPCT_EADMM', 'z1', 208), 'T_MPCT_ADMM_cs': ('MPCT_ADMM', 'z', 650)}


def _ref(name):
    from oracle import refs
    return refs.get(name)


def test_discretised_model_matches_reference_fixture(golden):
    """[A B] of the ZOH-discretised 3-mass system against examples/cl_in_C/main_cl_in_C.c:96 (15 decimals)."""
    sys = sysmodel.oscillating_masses_sys()
    AB = np.hstack([sys['A'], sys['B']])
    ref = np.array(golden['main_cl_in_C_AB']['AB'])
    assert AB.shape == ref.shape == (6, 8)
    assert np.max(np.abs(AB - ref)) < 1e-15 + 1e-15       # printed with %1.15f


def test_steady_state_reference_matches_fixture(golden):
    sys = sysmodel.oscillating_masses_sys()
    xr = sysmodel.steady_state(sys, 0.5 * np.ones(2))
    assert np.max(np.abs(xr - np.array(golden['main_cl_in_C_xr']['xr']))) < 1e-12   # main_cl_in_C.c:89


@pytest.mark.parametrize('name', list(GOLD))
def test_reference_template_reproduces_golden_vector(name, golden):
    """tests/test_<F>_<method>.m: gap.opt = ||z - z_opt||_inf <= 1e-4 and exit flag > 0 (spcies_tester.m:260-296)."""
    ref, spec, cfg = _ref(name)
    gold_name, field, k_expected = GOLD[name]
    st = cfg['status']
    r = cfg['param'].get('r', None) if ref.has_r else None
    u, k, e, sol = ref.solve(st['x'], st['xr'], st['ur'], r)
    z_opt = np.array(golden[gold_name]['z_opt'])
    assert e == 1
    assert k == k_expected                                  # iteration counts observed when the survey probed the reference
    assert np.max(np.abs(sol[field][:len(z_opt)] - z_opt)) <= 1e-4
    assert np.allclose(u, [0.8, 0.8], atol=1e-6)


@pytest.fixture(scope='module')
def hmpc_qp():
    """H, q, G, b of the harmonic-MPC problem recovered from its *definition* (tests/hmpc_definition.py: the cost and the
    equality residuals evaluated as functions and probed with unit vectors), at the reference test point."""
    import hmpc_definition as hd
    from spcies_b200 import configs
    cfg = configs.reference_test('HMPC_ADMM_split')
    st = cfg['status']
    H, q, G, b = hd.quadratic_from_definition(st['x'], st['xr'], st['ur'], cfg['sys'], cfg['param'])
    z, k, ok = hd.admm_twin(H, q, G, b, cfg['sys'], cfg['param'], rho=2.0)
    assert ok
    z2, _, ok2 = hd.admm_twin(H, q, G, b, cfg['sys'], cfg['param'], rho=11.0)      # the answer does not depend on the penalty
    assert ok2 and np.max(np.abs(z - z2)) <= 1e-7
    return dict(H=H, q=q, G=G, b=b, z=z, cfg=cfg)


def test_hmpc_ingredients_match_the_problem_definition(hmpc_qp):
    """The Hessian blocks H11..H33 and the equality matrix G restated in spcies_b200/formulations/HMPC.py
    (compute_HMPC_ADMM_split_ingredients.m:71-142) against the matrices recovered from the problem definition: a wrong block
    formula (harmonic sums, cross terms, terminal / steady-state rows) shows up here -- it would otherwise be invisible, because
    the CUDA solver and the instantiated reference template are both fed from HMPC.py."""
    from spcies_b200.formulations import HMPC
    from spcies_b200.gen_controller import make_recipe
    cfg = hmpc_qp['cfg']
    v = HMPC.compute_HMPC_ADMM_split_ingredients(make_recipe(cfg['sys'], cfg['param'], **cfg['kw']))
    assert v['H'].shape == hmpc_qp['H'].shape and v['G'].shape == hmpc_qp['G'].shape
    assert np.max(np.abs(v['H'] - hmpc_qp['H'])) <= 1e-9
    assert np.max(np.abs(v['G'] - hmpc_qp['G'])) <= 1e-12
    st = cfg['status']
    b = np.zeros(v['G'].shape[0])
    b[:v['n']] = -v['A'] @ st['x']
    assert np.max(np.abs(b - hmpc_qp['b'])) <= 1e-12
    # q of the C template (code_HMPC_ADMM_split_C.c:115-129): -Te xr - Q x0 at x_e, -Q x0 at x_c, -Se ur at u_e
    n, m, N = v['n'], v['m'], v['N']
    q = np.zeros(v['dim'])
    o = (N - 1) * (n + m) + m
    q[o:o + n] = -(v['Te'] @ st['xr'] + v['Q'] @ st['x'])
    q[o + 2 * n:o + 3 * n] = -(v['Q'] @ st['x'])
    q[o + 3 * n:o + 3 * n + m] = -(v['Se'] @ st['ur'])
    assert np.max(np.abs(q - hmpc_qp['q'])) <= 1e-9


@pytest.mark.parametrize('name', ['T_HMPC_ADMM_split', 'T_HMPC_SADMM_split', 'T_HMPC_ADMM'])
def test_hmpc_reference_against_independent_qp_solve(name, hmpc_qp):
    """The reference's HMPC golden vector is stale (8e-2 off what the current ingredients encode, for the
    split and the non-split solver alike), so the instantiated templates are pinned by an independent solve of the QP built from
    the problem definition (dense NumPy ADMM run to 1e-10, tests/hmpc_definition.py): ||z - z*||_inf <= 1e-4, the tolerance of
    the reference's own comparison (tests/spcies_tester.m:261); measured 1.7e-5 at the templates' exit tolerance 1e-7."""
    ref, spec, cfg = _ref(name)
    st = cfg['status']
    u, k, e, sol = ref.solve(st['x'], st['xr'], st['ur'])
    assert e == 1
    z = sol['z']
    assert np.max(np.abs(z - hmpc_qp['z'])) <= 1e-4
    assert np.max(np.abs(hmpc_qp['G'] @ z - hmpc_qp['b'])) <= 1e-5       # dynamics + harmonic steady-state equalities
    assert np.allclose(u, hmpc_qp['z'][:2], atol=1e-5)


def test_semiband_reference_against_the_mpct_golden_vector(golden):
    """MPCT_ADMM_semiband has no test of its own in the reference; it solves the same MPCT problem as MPCT_ADMM_cs, whose golden
    vector (tests/test_MPCT_ADMM.m:31, extended-state layout (x_j, x_s, u_j, u_s) per stage) therefore pins the restated semiband
    ingredients: predicted states and inputs, artificial reference."""
    ref, spec, cfg = _ref('T_MPCT_ADMM_semiband')
    st = cfg['status']
    u, k, e, sol = ref.solve(st['x'], st['xr'], st['ur'])
    assert e == 1
    n, m, N = spec.dims['n'], spec.dims['m'], spec.dims['N']
    z = sol['z'].reshape(N + 1, n + m)                       # (x_0, u_0), ..., (x_{N-1}, u_{N-1}), (x_s, u_s)
    zc = np.array(golden['MPCT_ADMM']['z_opt']).reshape(N, 2 * (n + m))
    assert np.max(np.abs(z[:N, :n] - zc[:, :n])) <= 1e-4 and np.max(np.abs(z[:N, n:] - zc[:, 2 * n:2 * n + m])) <= 1e-4
    assert np.max(np.abs(z[N, :n] - zc[0, n:2 * n])) <= 1e-4 and np.max(np.abs(z[N, n:] - zc[0, 2 * n + m:])) <= 1e-4
    assert np.allclose(u, [0.8, 0.8], atol=1e-6)


def test_semiband_qp_step_is_the_kkt_solve():
    """One iteration from zero of the instantiated semiband template == the KKT solve of [H + rho I, G'; G, 0] [z; mu] = [-q; b]:
    the identity the tensor-core engine of this solver is built on (csrc/MPCT_ADMM_semiband_mma.cuh)."""
    from oracle import instantiate
    from spcies_b200 import configs, make_spec
    cfg = configs.reference_test('MPCT_ADMM_semiband', k_max=1)
    spec = make_spec(cfg['sys'], cfg['param'], save_name='K1_MPCT_ADMM_semiband', **cfg['kw'])
    if not instantiate.reference_available():
        pytest.skip('needs the reference tree to instantiate a one-iteration solver')
    ref = instantiate.make_reference(spec, 'ref_K1_MPCT_ADMM_semiband')
    st = cfg['status']
    u, k, e, sol = ref.solve(st['x'], st['xr'], st['ur'])
    v = spec.vars
    H, G, n, m, N = v['H'], v['G'], v['n'], v['m'], v['N']
    L, LM = H.shape[0], G.shape[0]
    K = np.block([[H + v['rho'] * np.eye(L), G.T], [G, np.zeros((LM, LM))]])
    p = np.zeros(L)
    p[N * (n + m):N * (n + m) + n] = -v['T'] @ st['xr']
    p[N * (n + m) + n:] = -v['S'] @ st['ur']
    b = np.zeros(LM)
    b[:n] = st['x']
    z = np.linalg.solve(K, np.concatenate([-p, b]))[:L]
    assert k == 1 and np.max(np.abs(z - sol['z'])) <= 1e-10


def test_time_varying_reference_with_the_nominal_model_is_the_constant_model_reference():
    """The reference template compiled with `#define TIME_VARYING 1` (on-line block-Cholesky recursion, code_laxMPC_FISTA_C.c:139-262)
    and called with the nominal A, B, Q, R, bounds reproduces the constant-model template: same k, u_opt to rounding.  This pins the
    restated off-line ingredients (Alpha, Beta, QRi of spcies_b200/formulations/laxMPC.py) against the reference's own recursion."""
    from spcies_b200 import sysmodel
    ref, spec, cfg = _ref('TV_laxMPC_FISTA')
    r0 = _ref('C2_laxMPC_FISTA')[0]
    s = cfg['sys']
    B = 64
    b = sysmodel.synthetic_batch(s, B, seed=5)
    tv = (np.tile(s['A'], (B, 1, 1)), np.tile(s['B'], (B, 1, 1)), np.tile(np.diag(cfg['param']['Q']), (B, 1)),
          np.tile(np.diag(cfg['param']['R']), (B, 1)))
    LB = np.tile(np.concatenate([s['LBx'], s['LBu']]), (B, 1))
    UB = np.tile(np.concatenate([s['UBx'], s['UBu']]), (B, 1))
    u, k, e = ref.solve_batch(b['x0'], b['xr'], b['ur'], tv=tv, LB=LB, UB=UB, threads=4)
    u0, k0, e0 = r0.solve_batch(b['x0'], b['xr'], b['ur'])
    assert np.array_equal(k, k0) and np.array_equal(e, e0)
    assert np.max(np.abs(u - u0)) <= 1e-10
