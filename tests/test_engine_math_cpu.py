"""CPU: the algebra behind two engines, checked in NumPy without a GPU.

* latency engine / dense FISTA policy (csrc/MPC_FISTA_single.cuh, MPC_FISTA_dense.cuh): the FISTA iteration of
  code_laxMPC_FISTA_C.c:323-389 restated in the space of the primal variable -- mu+ = v + g + P z, v+ = mu+ + beta_k (mu+ - mu),
  z+ = clip(Hd o (q + v+)), P = E' W^-1 E, g = E' W^-1 b -- gives the iterates, the iteration count and the exit flag of the dual
  form (tests/fista_twin.py, the restatement of the reference's MATLAB twin).
* structured ADMM_soc engine (csrc/ellipMPC_ADMM_soc_band.cuh): once the two rows that pin t are taken out, W = Gh Hh^-1 Gh' of
  compute_ellipMPC_ADMM_soc_ingredients.m:131-147 is block tridiagonal with N + 1 blocks, and the block-wise evaluation of
  primal_hat (dynamics blocks, dense terminal block, cone rows) equals the reference's chain."""
import numpy as np

import fista_twin
from spcies_b200 import prebuilt, sysmodel


def _primal_form_solve(P, x0, xr, ur, tol, k_max):
    n, m, N = P['n'], P['m'], P['N']
    E = -P['Aeq']                                            # r = b + E z in the reference's sign convention
    b = np.zeros(N * n)
    b[:n] = -P['A'] @ x0
    q = -np.concatenate([P['R'] @ ur, np.tile(np.concatenate([P['Q'] @ xr, P['R'] @ ur]), N - 1), P['T'] @ xr])
    Pm, g = E.T @ P['Wi'] @ E, E.T @ P['Wi'] @ b
    # z(y) of the twin is clip(-(q - Aeq' y) / Hd) = clip(-(q + E' y) / Hd): with v = E' y
    zof = lambda v: np.clip(-(q + v) / P['Hd'], P['LB'], P['UB'])
    mu = np.zeros_like(q)
    v = np.zeros_like(q)
    z = zof(v)
    t, k = 1.0, 0
    while True:                                              # pass 0 = the initial step
        r = b + E @ z
        if k > 0:
            if np.max(np.abs(r)) <= tol:
                return z[:m].copy(), k, 1
            if k >= k_max:
                return z[:m].copy(), k, -1
        mnew = v + g + Pm @ z
        if k == 0:
            v = mnew
        else:
            t1 = t
            t = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t1 * t1))
            v = mnew + (t1 - 1.0) / t * (mnew - mu)
        mu = mnew
        z = zof(v)
        k += 1


def test_fista_in_primal_space_is_the_dual_iteration():
    spec, cfg = prebuilt.spec_for('C2_laxMPC_FISTA')
    P = fista_twin.build(cfg['sys'], cfg['param'])
    tol, kmax = float(spec.define('tol')), int(spec.define('k_max'))
    b = sysmodel.synthetic_batch(cfg['sys'], 40, seed=3)
    for i in range(40):
        u0, k0, e0, _ = fista_twin.solve(P, b['x0'][i], b['xr'][i], b['ur'][i], tol=tol, k_max=kmax)
        u1, k1, e1 = _primal_form_solve(P, b['x0'][i], b['xr'][i], b['ur'][i], tol, kmax)
        assert e0 == e1 and abs(k0 - k1) <= 1
        if k0 == k1:
            assert np.max(np.abs(u0 - u1)) <= 1e-9


def test_soc_system_is_block_tridiagonal_without_the_rows_of_t():
    spec, cfg = prebuilt.spec_for('C4_ellipMPC_ADMM_soc')
    v = spec.vars
    n, m, N = v['n'], v['m'], v['N']
    nm = n + m
    Gh, Hh, W = v['Gh'], v['Hh'], v['W']
    Hhi = np.linalg.inv(Hh)
    NP, NR, DIM, NEQ = Hh.shape[0], Gh.shape[0], v['dim'], v['n_eq']
    rows = list(range(N * n)) + list(range(NEQ + 1, NEQ + 1 + n))
    Wc = W[np.ix_(rows, rows)]
    assert np.abs(W[np.ix_(rows, [N * n, NEQ])]).max() <= 1e-13              # the t rows decouple
    R = np.linalg.cholesky(Wc).T
    for i in range(R.shape[0]):
        assert np.abs(R[i, (i // n + 2) * n:]).max(initial=0.0) <= 1e-12      # block bidiagonal factor
    # block-wise evaluation of primal_hat against the chain  -Hhi q_hat - Hhi Gh' W^-1 (-Gh Hhi q_hat - bh)
    rng = np.random.default_rng(0)
    qh, x0, xr, r = rng.standard_normal(NP), rng.standard_normal(n), rng.standard_normal(n), 1.3
    bh = np.zeros(NR)
    bh[:n] = -v['A'] @ x0
    bh[NEQ - 1] = r
    bh[NEQ + 1:NEQ + 1 + n] = -v['PhiP'] @ xr
    nu = np.linalg.solve(W, -Gh @ Hhi @ qh - bh)
    ref = -Hhi @ qh - Hhi @ Gh.T @ nu
    zx = lambda blk: m + (blk - 1) * nm
    AB = Gh[n:2 * n, zx(1):zx(1) + nm]
    Hi = np.diag(Hhi)[zx(1):zx(1) + nm]
    HiN = Hhi[zx(N):zx(N) + n, zx(N):zx(N) + n]
    Ph = -Gh[NEQ + 1:NEQ + 1 + n, zx(N):zx(N) + n]
    s0 = np.zeros(nm)
    s0[:n], s0[n:] = -x0, Hi[n:] * qh[:m]
    s = [s0] + [Hi * qh[zx(blk):zx(blk) + nm] for blk in range(1, N)]
    qN = qh[zx(N):zx(N) + n]
    sN = HiN @ qN
    rr = [(s[blk + 1][:n] if blk + 1 < N else sN) - AB @ s[blk] for blk in range(N)]
    qs = qh[DIM + 1:DIM + 1 + n]
    rr.append(Ph @ sN - qs / v['rho'] + v['PhiP'] @ xr)
    nub = np.linalg.solve(Wc, np.concatenate(rr)).reshape(N + 1, n)
    out = np.zeros(NP)
    for blk in range(N):
        qb = np.zeros(nm)
        if blk == 0:
            qb[n:] = qh[:m]
        else:
            qb = qh[zx(blk):zx(blk) + nm].copy()
        c = qb.copy()
        if blk > 0:
            c[:n] -= nub[blk - 1]
        z = -Hi * (c + AB.T @ nub[blk])
        if blk == 0:
            out[:m] = z[n:]
        else:
            out[zx(blk):zx(blk) + nm] = z
    out[zx(N):zx(N) + n] = -HiN @ (qN - nub[N - 1] - Ph.T @ nub[N])
    out[DIM - 1] = out[DIM] = r                                             # t_hat = s_hat_0 = r
    out[DIM + 1:DIM + 1 + n] = -(qs + nub[N]) / v['rho']
    assert np.max(np.abs(out - ref)) <= 1e-10 * max(1.0, np.abs(ref).max())
