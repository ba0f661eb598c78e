"""GPU parity of the solver kernels against the instantiated reference C solver (oracle/_ref)."""
import numpy as np
import pytest

from spcies_b200 import prebuilt, sysmodel
from spcies_b200.solver import ARITH_EXACT, ARITH_FAST, TAIL_SINGLE, TAIL_TWO_PHASE

pytestmark = pytest.mark.gpu

# generated solver -> (golden vector in tests/golden/reference_golden.json, sol field compared with it)
# (None: the reference's HMPC golden vector is stale -- SURVEY.md section 4 -- so HMPC is pinned by the instantiated C only)
GOLD = {'T_laxMPC_FISTA': ('laxMPC_FISTA', 'z'), 'T_equMPC_FISTA': ('equMPC_FISTA', 'z'),
        'T_laxMPC_ADMM': ('laxMPC_ADMM', 'z'), 'T_equMPC_ADMM': ('equMPC_ADMM', 'z'),
        'T_ellipMPC_ADMM': ('ellipMPC_ADMM', 'z'), 'T_ellipMPC_ADMM_soc': ('ellipMPC_ADMM_soc', 'z'),
        'T_MPCT_EADMM': ('MPCT_EADMM', 'z1'), 'T_HMPC_ADMM_split': None, 'T_HMPC_SADMM_split': None,
        'T_HMPC_SADMM_split_sparse': None}      # sparse = true: the KKT L D L' branch (code_HMPC_ADMM_split_C.c:193-209)
TEST_SOLVERS = list(GOLD)
BATCH_SOLVERS = TEST_SOLVERS + ['C2_laxMPC_FISTA', 'C3_equMPC_ADMM', 'C4_ellipMPC_ADMM_soc', 'T_equMPC_ADMM_vb']   # _vb: per-stage bounds
LONG_HORIZON = ['C5a_HMPC_SADMM_split', 'C5b_MPCT_EADMM']      # N = 50: per-instance state in the global scratch


def _batch(sol, cfg, B, seed):
    b = sysmodel.synthetic_batch(cfg['sys'], B, seed=seed, with_r=sol.has_r)
    return b, (dict(r=b['r']) if sol.has_r else {})


def _cmp_fields(spec, s, sr, exact):
    """Compare every debug vector of sol_<name>.  MPCT_EADMM: the reference's DEBUG copy of lambda only fills the
    first nn_ of every nm_ entries, packed (code_MPCT_EADMM_C.c:509-513) -- compare those."""
    for f, _len in spec.sol_fields:
        a, b = s[f], sr[f]
        if spec.func_name == 'MPCT_EADMM' and f == 'lambda':
            n, nm = spec.dims['n'], spec.dims['n'] + spec.dims['m']
            nblk = a.shape[-1] // nm
            a = a.reshape(a.shape[:-1] + (nblk, nm))[..., :n].reshape(a.shape[:-1] + (nblk * n,))
            b = b[..., :nblk * n]
        if exact:
            assert np.array_equal(a, b), f
        else:
            assert np.max(np.abs(a - b)) <= 1e-9, f


def _ref(name):
    from oracle import refs
    return refs.get(name)[0]


from _parity import abs_err as _abs_err, rel_err as _rel_err   # true element-wise relative error (floor 1e-3), tests/_parity.py


@pytest.mark.parametrize('name', TEST_SOLVERS)
def test_single_instance_symbol_vs_golden_and_reference(name, golden):
    """Reference test point (tests/spcies_tester.m:114-116) through the UNCHANGED single-instance symbol."""
    sol, spec, cfg = prebuilt.get(name)
    st = cfg['status']
    r = cfg['param'].get('r', None) if sol.has_r else None
    u, k, e, s = sol.solve(st['x'], st['xr'], st['ur'], r)
    ur_, kr, er, sr = _ref(name).solve(st['x'], st['xr'], st['ur'], r)
    assert e == er == 1
    assert abs(k - kr) <= 1
    if GOLD[name] is not None:
        gold_name, field = GOLD[name]
        z_opt = np.array(golden[gold_name]['z_opt'])
        z = s[field][:len(z_opt)]                            # ADMM_soc compares z(1:end-1), test_ellipMPC_ADMM_soc.m:46
        assert np.max(np.abs(z - z_opt)) <= 1e-4             # tol_opt of tests/spcies_tester.m:261
    _cmp_fields(spec, s, sr, exact=False)
    assert _rel_err(u, ur_) <= 1e-9


@pytest.mark.parametrize('name', BATCH_SOLVERS)
def test_exact_mode_is_bit_identical(name):
    """ARITH_EXACT: same IEEE operations in the same order as gcc -O3 => identical bits, k and e_flag."""
    sol, spec, cfg = prebuilt.get(name)
    batch, kw = _batch(sol, cfg, 4096 if 'FISTA' in name else 512, seed=1)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, **kw)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8, **kw)
    assert np.array_equal(e, er)
    assert np.array_equal(k, kr)
    assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    assert info['sum_k'] == int(k.sum())
    assert info['n_not_converged'] == int((e == -1).sum())


@pytest.mark.parametrize('name', BATCH_SOLVERS)
def test_fast_mode_parity(name):
    """ARITH_FAST (FMA): u_opt within 1e-9 relative, e_flag identical, |dk| <= 1 (BASELINE.json north_star)."""
    sol, spec, cfg = prebuilt.get(name)
    batch, kw = _batch(sol, cfg, 8192 if 'FISTA' in name else 512, seed=2)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, **kw)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8, **kw)
    assert np.array_equal(e, er)
    assert np.max(np.abs(k - kr)) <= 1
    same = k == kr
    conv = er == 1
    assert _rel_err(u[same & conv], ur_[same & conv]) <= 1e-9
    # instances that hit k_max return an iterate that is not a solution (diverging duals amplify rounding): 1e-7
    assert _abs_err(u[same & ~conv], ur_[same & ~conv]) <= 1e-7
    # an instance whose k moved by one stops one iterate earlier/later: compare at the solver tolerance
    if (~same).any():
        assert _abs_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol', spec.define('tol_p')))


def test_float_precision_parity():
    """precision = 'float' (BASELINE.json configs[2], equMPC ADMM N = 20).  The reference generated with precision = 'float' keeps
    double arithmetic on float-rounded constants (platforms/+C_code/dec_var.m:16-17 only changes the declared type of the
    constants); so does the CUDA solver by default (constants float, arithmetic double, tensor-core engine): the north-star gate
    holds as for double -- e_flag identical, |dk| <= 1, u_opt within 1e-5 (measured: 1e-9)."""
    from _parity import gate
    sol, spec, cfg = prebuilt.get('C3f_equMPC_ADMM')
    assert sol.precision == 'float' and sol.arithmetic == 'double'
    batch, kw = _batch(sol, cfg, 16384, seed=41)
    ur_, kr, er = _ref('C3f_equMPC_ADMM').solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST)
    gate(spec, u, k, e, ur_, kr, er, tol=1e-9)
    assert info['sum_k'] == int(k.sum())
    u, k, e, info = sol.solve_batch(batch['x0'][:2048], batch['xr'][:2048], batch['ur'][:2048], arith=ARITH_EXACT)
    assert np.array_equal(k, kr[:2048]) and np.array_equal(e, er[:2048])
    assert np.array_equal(u.view(np.uint64), ur_[:2048].view(np.uint64))        # bit-identical to the float-generated reference


def test_float_arithmetic_parity():
    """options.solver['float_arithmetic']: true single-precision kernels.  Against the float-generated reference (double
    arithmetic): e_flag identical, u_opt within 1e-5 of the input range where the iteration counts agree, |dk| <= 1 on all but a
    handful of instances (float rounding of the residuals moves the exit test; measured 4 of 16384 with |dk| up to 5), u_opt
    within 10 tol there."""
    sol, spec, cfg = prebuilt.get('C3ff_equMPC_ADMM')
    assert sol.precision == 'float' and sol.arithmetic == 'float'
    batch, kw = _batch(sol, cfg, 16384, seed=41)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST)
    ur_, kr, er = _ref('C3ff_equMPC_ADMM').solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16)
    assert np.array_equal(e, er)
    dk = np.abs(k - kr)
    assert (dk > 1).mean() <= 1e-3 and dk.max() <= 8
    same = (dk == 0) & (er == 1)
    # float: relative to the input range (|u| <= 0.8): a float iterate cannot be 1e-5 relative to an entry that is itself 1e-3
    assert _rel_err(u[same], ur_[same], floor=0.8) <= 1e-5
    assert _abs_err(u[~same], ur_[~same]) <= 10 * float(spec.define('tol'))
    assert info['sum_k'] == int(k.sum())


@pytest.mark.parametrize('name', TEST_SOLVERS)
def test_debug_payload_batch(name):
    """sol_<name> payload (DEBUG builds) of a ragged batch, exact mode: every vector bit-identical."""
    sol, spec, cfg = prebuilt.get(name)
    batch, kw = _batch(sol, cfg, 257 if 'HMPC' not in name else 65, seed=3)
    u, k, e, info, s = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, want_sol=True, **kw)
    ur_, kr, er, sr = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], want_sol=True, threads=8, **kw)
    assert np.array_equal(k, kr) and np.array_equal(e, er)
    _cmp_fields(spec, s, sr, exact=True)


@pytest.mark.parametrize('name', LONG_HORIZON)
def test_long_horizon_exact(name):
    """N = 50 configurations of BASELINE.json configs[4] (state in the L2-resident global scratch), exact mode."""
    sol, spec, cfg = prebuilt.get(name)
    batch, kw = _batch(sol, cfg, 192 if 'HMPC' in name else 640, seed=8)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, **kw)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=16, **kw)
    assert np.array_equal(e, er) and np.array_equal(k, kr)
    assert np.array_equal(u, ur_)


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


@pytest.mark.parametrize('name', ['T_laxMPC_FISTA', 'T_equMPC_FISTA', 'C2_laxMPC_FISTA'])
@pytest.mark.parametrize('grace', [1, 7, 40])
def test_tail_park_and_resume_is_invisible(name, grace):
    """Tail handling (two launches: park the instances still iterating `grace` iterations after the queue ran dry,
    resume them in a second launch): the iterates are copied verbatim, so exact mode stays bit-identical to the
    reference and to the single-launch run, whatever the parking point."""
    sol, spec, cfg = prebuilt.get(name)
    batch, kw = _batch(sol, cfg, 3000, seed=11)
    ur_, kr, er = _ref(name).solve_batch(batch['x0'], batch['xr'], batch['ur'], threads=8, **kw)
    u1, k1, e1, i1 = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, tail_mode=TAIL_SINGLE)
    u2, k2, e2, i2 = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, tail_mode=TAIL_TWO_PHASE,
                                     tail_grace=grace)
    assert i1['launches'] == 1 and i1['parked'] == 0
    assert i2['launches'] == 2 and i2['parked'] > 0
    for u, k, e in ((u1, k1, e1), (u2, k2, e2)):
        assert np.array_equal(e, er) and np.array_equal(k, kr)
        assert np.array_equal(u.view(np.uint64), ur_.view(np.uint64))
    assert i2['sum_k'] == int(kr.sum()) and i2['n_not_converged'] == int((er == -1).sum())
    # fast arithmetic through the same path
    u3, k3, e3, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_FAST, tail_mode=TAIL_TWO_PHASE,
                                    tail_grace=grace)
    assert np.array_equal(e3, er) and np.max(np.abs(k3 - kr)) <= 1
    same = k3 == kr
    conv = er == 1
    assert _rel_err(u3[same & conv], ur_[same & conv]) <= 1e-9
    assert _abs_err(u3[same & ~conv], ur_[same & ~conv]) <= 1e-7


def test_tail_park_and_resume_with_per_instance_bounds():
    """Same with opts.LB / UB (the VARB kernels) and the debug payload: two-phase == single launch, bit for bit."""
    sol, spec, cfg = prebuilt.get('T_laxMPC_FISTA')
    B = 1500
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=12)
    rng = np.random.default_rng(5)
    nm = sol.n + sol.m
    LB = np.tile(np.concatenate([cfg['sys']['LBx'], cfg['sys']['LBu']]), (B, 1))
    UB = np.tile(np.concatenate([cfg['sys']['UBx'], cfg['sys']['UBu']]), (B, 1))
    UB[:, :3] = rng.uniform(0.25, 0.35, size=(B, 3))
    LB[:, sol.n:] = -rng.uniform(0.5, 0.8, size=(B, sol.m))
    assert LB.shape == (B, nm)
    a = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=LB, UB=UB, arith=ARITH_EXACT, tail_mode=TAIL_SINGLE,
                        want_sol=True)
    b = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], LB=LB, UB=UB, arith=ARITH_EXACT, tail_mode=TAIL_TWO_PHASE,
                        tail_grace=5, want_sol=True)
    assert b[3]['parked'] > 0
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for f, _len in spec.sol_fields:
        assert np.array_equal(a[4][f], b[4][f]), f


def test_large_host_batch_pipelined_copies_and_auto_tail():
    """A host-buffer call large enough for the pipelined host->device copies (chunks arrive under the running kernel,
    the queue waits on the watermark) and for the automatic two-launch tail handling: exact mode, checked against the
    reference on the first / last / a random subset of the instances, and against a device-resident single-launch run
    of the whole batch through the checksum of all outputs."""
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    B = 330_000
    batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=21)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT)
    assert info['launches'] == 2 and info['parked'] > 0
    rng = np.random.default_rng(3)
    idx = np.unique(np.concatenate([np.arange(3000), np.arange(B - 3000, B), rng.integers(0, B, 6000)]))
    ur_, kr, er = _ref('C2_laxMPC_FISTA').solve_batch(batch['x0'][idx], batch['xr'][idx], batch['ur'][idx], threads=16)
    assert np.array_equal(k[idx], kr) and np.array_equal(e[idx], er)
    assert np.array_equal(u[idx].view(np.uint64), ur_.view(np.uint64))
    u1, k1, e1, info1 = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], arith=ARITH_EXACT, tail_mode=TAIL_SINGLE)
    assert info1['launches'] == 1
    assert np.array_equal(u.view(np.uint64), u1.view(np.uint64)) and np.array_equal(k, k1) and np.array_equal(e, e1)
    assert info['sum_k'] == int(k.sum()) == info1['sum_k']


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


def test_plain_c_harness_runs():
    """harness/main_batch.c: the reference's plain-C caller pattern (examples/cl_in_C/main_cl_in_C.c) on the CUDA library."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, 'harness', 'main_batch')
    assert os.path.exists(exe), 'run __graft_entry__.build() first'
    out = subprocess.run([exe, '4096', '2'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert 'single solve: u = [0.800000 0.800000]' in out.stdout and 'e_flag = 1' in out.stdout
    assert 'closed loop step 2' in out.stdout


@pytest.mark.parametrize('name,B', [('C3_equMPC_ADMM', 330_000), ('C4_ellipMPC_ADMM_soc', 300_000), ('C5b_MPCT_EADMM', 90_000),
                                    ('C5a_HMPC_SADMM_split', 80_000)])
def test_pipelined_host_copies_every_engine(name, B):
    """Host-buffer calls large enough for the pipelined host->device copies (chunks arrive under the running kernel, the
    instance queue waits on the watermark): every instance is solved exactly once and gets the bits of a call on the same
    instances made in pieces below the pipelining threshold (results do not depend on which lane group solves an instance)."""
    sol, spec, cfg = prebuilt.get(name)
    batch, kw = _batch(sol, cfg, B, seed=91)
    u, k, e, info = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], **kw)
    assert info['sum_k'] == int(k.sum()) and info['n_not_converged'] == int((e == -1).sum())
    assert set(np.unique(e)) <= {1, -1} and k.min() >= 1
    rng = np.random.default_rng(7)
    for lo in (0, B - 20_000, int(rng.integers(20_000, B - 40_000))):
        sl = slice(lo, lo + 20_000)
        kws = {kk: v[sl] for kk, v in kw.items()}
        u2, k2, e2, _ = sol.solve_batch(batch['x0'][sl], batch['xr'][sl], batch['ur'][sl], **kws)
        assert np.array_equal(u[sl].view(np.uint64), u2.view(np.uint64)) and np.array_equal(k[sl], k2) and np.array_equal(e[sl], e2)


def test_results_written_directly_into_pinned_host_arrays():
    """When the caller's u_opt / k / e_flag arrays are pinned host memory the kernel writes them directly (no device->host
    copies); the results are the bits of the call with pageable arrays."""
    import torch
    sol, spec, cfg = prebuilt.get('C2_laxMPC_FISTA')
    for B in (200, 40_000):
        batch = sysmodel.synthetic_batch(cfg['sys'], B, seed=92)
        u0, k0, e0, _ = sol.solve_batch(batch['x0'], batch['xr'], batch['ur'])
        hu = torch.full((B, sol.m), float('nan'), dtype=torch.float64).pin_memory()
        hk = torch.full((B,), -7, dtype=torch.int32).pin_memory()
        he = torch.full((B,), -7, dtype=torch.int32).pin_memory()
        sol.solve_batch(batch['x0'], batch['xr'], batch['ur'], out=(hu.numpy(), hk.numpy(), he.numpy()))
        assert np.array_equal(hu.numpy().view(np.uint64), u0.view(np.uint64))
        assert np.array_equal(hk.numpy(), k0) and np.array_equal(he.numpy(), e0)
