"""MPCT (MPC for tracking with artificial reference) -- EADMM recipe.

Host-side restatement of formulations/+MPCT/compute_MPCT_EADMM_ingredients.m:21-316
and cons_MPCT_EADMM_C.m.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .common import (Row, SolverSpec, alpha_beta_from_chol, chol_upper, default_defines,
                     engineering_rows, get_sys_param, isdiag, scaling_vars, var_options)


def compute_MPCT_EADMM_ingredients(recipe):
    A, B, n, m, N = get_sys_param(recipe)
    sys, param, opt = recipe.sys, recipe.param, recipe.options
    nm = n + m
    Q, R, T, S = (np.asarray(param[k], float) for k in ('Q', 'R', 'T', 'S'))
    inf = float(opt.inf_value)
    LBx = np.asarray(sys.get('LBx', -inf * np.ones(n)), float).ravel()
    UBx = np.asarray(sys.get('UBx', inf * np.ones(n)), float).ravel()
    LBu = np.asarray(sys.get('LBu', -inf * np.ones(m)), float).ravel()
    UBu = np.asarray(sys.get('UBu', inf * np.ones(m)), float).ravel()

    solver = opt.solver
    if 'rho' in solver:                         # compute_MPCT_EADMM_ingredients.m:70-73
        rho_base, rho_mult = float(solver['rho']), 1.0
    else:
        rho_base, rho_mult = float(solver['rho_base']), float(solver['rho_mult'])

    # Penalty vector; MATLAB end-relative slices restated 0-based
    L = (N + 1) * nm + n + nm
    rho = rho_base * np.ones(L)
    big = rho_mult * rho_base
    rho[0:n] = big                              # x_0 = x                      (6b)
    rho[n:2 * n] = big                          # initial z1 + z2 + z3 = 0     (6i), i = 0
    rho[L - 2 * nm:L - nm - m] = big            # final   z1 + z2 + z3 = 0     (6i), i = N
    rho[L - nm:L - m] = big                     # x_N = x_s                    (6k)
    rho[L - 2 * nm + n:L - nm] = big            # final   z1 + z2 + z3 = 0     (6j), i = N
    rho[L - m:L] = big                          # u_N = u_s                    (6l)

    A1 = -np.vstack([np.hstack([-np.eye(n), np.zeros((n, N * nm + m))]),
                     np.eye((N + 1) * nm),
                     np.hstack([np.zeros((nm, N * nm)), np.eye(nm)])])
    A2 = np.vstack([np.zeros((n, nm)), np.kron(np.ones((N, 1)), np.eye(nm)),
                    np.kron(np.ones((2, 1)), np.eye(nm))])
    A3 = np.vstack([np.zeros((n, (N + 1) * nm)), np.eye((N + 1) * nm), np.zeros((nm, (N + 1) * nm))])

    rA1 = rho[:, None] * A1                     # ``rho.*A`` scales rows
    H1 = rA1.T @ A1
    H1i = 1.0 / np.diag(H1)

    H2 = sla.block_diag(T, S) + (rho[:, None] * A2).T @ A2
    Az2 = np.hstack([A - np.eye(n), B])
    H2i = np.linalg.inv(H2)
    W2 = H2i @ Az2.T @ np.linalg.inv(Az2 @ H2i @ Az2.T) @ Az2 @ H2i - H2i

    H3 = np.kron(np.eye(N + 1), sla.block_diag(Q, R)) + (rho[:, None] * A3).T @ A3
    Az3 = np.kron(np.eye(N), np.hstack([A, B]))
    for j in range(N - 1):                      # -I blocks, no growth here (Az3 is N blocks wide)
        c0 = j * nm + nm
        Az3[j * n:(j + 1) * n, c0:c0 + n] = -np.eye(n)
    Az3 = np.hstack([Az3, np.vstack([np.zeros(((N - 1) * n, n)), -np.eye(n)]), np.zeros((N * n, m))])
    H3inv = np.linalg.inv(H3)
    W3 = Az3 @ H3inv @ Az3.T
    W3c = chol_upper(W3)

    diag_ok = bool(opt.force_diagonal) and isdiag(Q) and isdiag(R)

    v = dict(n=n, m=m, N=N, force_diagonal=diag_ok)
    v['H1i'] = H1i.reshape(N + 1, nm).copy()
    if diag_ok:
        v['H3i'] = (1.0 / np.diag(H3)).reshape(N + 1, nm).copy()
    else:
        Qm = np.linalg.inv(Q + big * np.eye(n))
        Rm = np.linalg.inv(R + big * np.eye(m))
        Qb = np.linalg.inv(Q + rho_base * np.eye(n))
        Rb = np.linalg.inv(R + rho_base * np.eye(m))
        v.update(Q_mult_inv=Qm, Q_base_inv=Qb, R_mult_inv=Rm, R_base_inv=Rb,
                 AB_base_inv=np.hstack([A, B]) @ sla.block_diag(Qb, Rb),
                 AB_mult_inv=np.hstack([A, B]) @ sla.block_diag(Qm, Rb))
    v['AB'] = np.hstack([A, B])
    v['W2'] = W2
    v['T'] = -T
    v['S'] = -S

    def clipinf(x, s):
        x = np.array(x, float)
        x[np.isinf(x)] = s * inf
        return x
    eps_x, eps_u = float(solver['epsilon_x']), float(solver['epsilon_u'])
    v['LB'] = clipinf(np.concatenate([LBx, LBu]), -1)
    v['UB'] = clipinf(np.concatenate([UBx, UBu]), +1)
    v['LBs'] = clipinf(np.concatenate([LBx + eps_x, LBu + eps_u]), -1)
    v['UBs'] = clipinf(np.concatenate([UBx - eps_x, UBu - eps_u]), +1)
    v['LB0'] = np.concatenate([-inf * np.ones(n), LBu])
    v['UB0'] = np.concatenate([inf * np.ones(n), UBu])
    v['rho'] = rho[n:L - nm].reshape(N + 1, nm).copy()
    v['rho_0'] = np.concatenate([rho[:n], np.zeros(m)])
    v['rho_s'] = rho[L - nm:].copy()
    v['Alpha'], v['Beta'] = alpha_beta_from_chol(W3c, n, N)
    v.update(scaling_vars(sys, n, m))
    # dense pieces for the non-sparse twin (vars_nonsparse in the reference)
    v['dense'] = dict(A1=A1, A2=A2, A3=A3, rho=rho, H1i=H1i, T=T, S=S, W2=W2, H3inv=H3inv, Az3=Az3, W3=W3,
                      LB=np.concatenate([v['LB0'], np.kron(np.ones(N - 1), v['LB']), v['LBs']]),
                      UB=np.concatenate([v['UB0'], np.kron(np.ones(N - 1), v['UB']), v['UBs']]))
    return v


def cons_MPCT_EADMM(recipe) -> SolverSpec:
    opts = recipe.options
    v = compute_MPCT_EADMM_ingredients(recipe)
    # cons_MPCT_EADMM_C.m: force_diagonal follows whether H3i exists *before* the default #defines are built
    opts = opts.copy()
    opts.force_diagonal = bool(v['force_diagonal'])
    n, m, N = v['n'], v['m'], v['N']
    vopt = var_options(opts)
    prec = opts.precision
    D = ('define',)
    defs = default_defines(opts)
    defs += [Row('nn_', n, True, 'uint', D), Row('mm_', m, True, 'uint', D),
             Row('nm_', n + m, True, 'uint', D), Row('NN_', N, True, 'uint', D),
             Row('k_max', int(opts.solver['k_max']), True, 'uint', D),
             Row('tol', float(opts.solver['tol']), True, 'float', D)]
    names = [('rho', 'rho'), ('rho_0', 'rho_0'), ('rho_s', 'rho_s'), ('LB', 'LB'), ('UB', 'UB'),
             ('LB_0', 'LB0'), ('UB_0', 'UB0'), ('LB_s', 'LBs'), ('UB_s', 'UBs'), ('AB', 'AB'),
             ('T', 'T'), ('S', 'S'), ('Alpha', 'Alpha'), ('Beta', 'Beta'), ('H1i', 'H1i'), ('W2', 'W2')]
    consts = [Row(cn, v[vn], True, prec, vopt) for cn, vn in names]
    if opts.force_diagonal:
        consts.append(Row('H3i', v['H3i'], True, prec, vopt))
    else:
        consts += [Row('Q_bi', v['Q_base_inv'], True, prec, vopt), Row('Q_mi', v['Q_mult_inv'], True, prec, vopt),
                   Row('R_bi', v['R_base_inv'], True, prec, vopt), Row('R_mi', v['R_mult_inv'], True, prec, vopt),
                   Row('AB_bi', v['AB_base_inv'], True, prec, vopt), Row('AB_mi', v['AB_mult_inv'], True, prec, vopt)]
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    nm = n + m
    return SolverSpec(
        formulation='MPCT', method='EADMM', submethod='', func_name='MPCT_EADMM', kernel='MPCT_EADMM',
        defines=defs, constants=consts,
        ref_code='formulations/+MPCT/code_MPCT_EADMM_C.c',
        ref_header='formulations/+MPCT/header_MPCT_EADMM_C.h',
        sol_fields=(('z1', (N + 1) * nm), ('z2', nm), ('z3', (N + 1) * nm), ('lambda', (N + 3) * nm)),
        vars=v, dims=dict(n=n, m=m, N=N))


# --------------------------------------------------------------------------------------
# ADMM on the extended state space (submethod 'cs', the default of MPCT / ADMM)
# --------------------------------------------------------------------------------------
def compute_MPCT_ADMM_cs_ingredients(recipe):
    """formulations/+MPCT/compute_MPCT_ADMM_cs_ingredients.m:18-195: the state / input are extended with the artificial
    reference, z_j = (x_j, x_s), v_j = (u_j, u_s); the equality-constrained QP of the z update is solved through the sparse
    chain  -A Hi (CSR)  ->  L D L' of W = A Hi A' (CSC)  ->  -Hi, -Hi A' (CSR)."""
    from .. import sp_utils
    A, B, n, m, N = get_sys_param(recipe)
    sys, param, opt = recipe.sys, recipe.param, recipe.options
    nm = n + m
    Q, R, T, S = (np.asarray(param[k], float) for k in ('Q', 'R', 'T', 'S'))
    inf = float(opt.inf_value)
    LBx = np.asarray(sys.get('LBx', -inf * np.ones(n)), float).ravel()
    UBx = np.asarray(sys.get('UBx', inf * np.ones(n)), float).ravel()
    LBu = np.asarray(sys.get('LBu', -inf * np.ones(m)), float).ravel()
    UBu = np.asarray(sys.get('UBu', inf * np.ones(m)), float).ravel()
    solver = opt.solver
    rho = solver['rho']
    if np.isscalar(rho) and solver.get('force_vector_rho', False):
        rho = float(rho) * np.ones(N * nm)          # (sic) N (n+m) entries, :71 -- half the 2 N (n+m) of the decision vector
    scalar = bool(np.isscalar(rho))

    Qz = np.block([[Q, -Q], [-Q, Q + T / N]])
    Rz = np.block([[R, -R], [-R, R + S / N]])
    H = np.kron(np.eye(N), sla.block_diag(Qz, Rz))
    Hhat = H + (float(rho) * np.eye(2 * N * nm) if scalar else np.diag(np.asarray(rho, float)))

    Zn, Zm = np.zeros((n, n)), np.zeros((m, m))
    AA = np.vstack([np.hstack([A, Zn]), np.hstack([Zn, np.eye(n)]), np.zeros((m, 2 * n))])
    BB = np.vstack([np.hstack([B, np.zeros((n, m))]), np.zeros((n, 2 * m)), np.hstack([Zm, np.eye(m)])])
    II = np.vstack([np.hstack([-np.eye(n), np.zeros((n, n + 2 * m))]),
                    np.hstack([Zn, -np.eye(n), np.zeros((n, 2 * m))]),
                    np.hstack([np.zeros((m, 2 * n + m)), -np.eye(m)])])
    dn = 2 * n + m                                  # rows per stage
    dnm = 2 * nm                                    # columns per stage
    Aeq = sla.block_diag(np.kron(np.eye(N - 1), np.hstack([AA, BB])), np.hstack([A, -np.eye(n), B, np.zeros((n, m))]))
    for j in range(1, N):                           # the II blocks, one block-column to the right of each [AA BB]
        r0 = (j - 1) * dn
        Aeq[r0:r0 + dn, j * dnm:(j + 1) * dnm] = II
    init_cond = np.vstack([np.hstack([np.eye(n), np.zeros((n, n + 2 * m))]),
                           np.hstack([Zn, A - np.eye(n), np.zeros((n, m)), B])])
    Aeq = np.vstack([np.hstack([init_cond, np.zeros((2 * n, Aeq.shape[1] - dnm))]), Aeq])

    eps_x, eps_u = float(solver['epsilon_x']), float(solver['epsilon_u'])
    LBz = np.concatenate([LBx, LBx + eps_x])
    UBz = np.concatenate([UBx, UBx - eps_x])
    LBv = np.concatenate([LBu, LBu + eps_u])
    UBv = np.concatenate([UBu, UBu - eps_u])
    LB = np.kron(np.ones(N), np.concatenate([LBz, LBv]))
    UB = np.kron(np.ones(N), np.concatenate([UBz, UBv]))

    Hinv = np.linalg.inv(Hhat)
    W = Aeq @ Hinv @ Aeq.T
    L_val, L_row, L_col, Dinv = sp_utils.full2LDL(W, for_LDLsolve=True)
    v = dict(n=n, m=m, N=N, rho_is_scalar=scalar)
    v['Tz'] = -(1.0 / N) * T
    v['Sz'] = -(1.0 / N) * S
    v['LB'], v['UB'] = LB, UB
    v['rho'] = float(rho) if scalar else np.asarray(rho, float)
    v['rho_i'] = 1.0 / v['rho']
    v['L_CSC'] = sp_utils.Sparse(val=L_val, row=L_row, col=L_col, nnz=len(L_val), nrow=W.shape[0], ncol=W.shape[0])
    v['Dinv'] = Dinv
    v['AHi_CSR'] = sp_utils.full2CSR(-Aeq @ Hinv)
    v['HiA_CSR'] = sp_utils.full2CSR(-Hinv @ Aeq.T)
    v['Hi_CSR'] = sp_utils.full2CSR(-Hinv)
    v.update(scaling_vars(sys, n, m))
    v['dense'] = dict(Hhat=Hhat, H=H, Aeq=Aeq, W=W, Hinv=Hinv)
    return v


def cons_MPCT_ADMM_cs(recipe) -> SolverSpec:
    """formulations/+MPCT/cons_MPCT_ADMM_cs_C.m:36-131."""
    opts = recipe.options
    v = compute_MPCT_ADMM_cs_ingredients(recipe)
    n, m, N = v['n'], v['m'], v['N']
    vopt = var_options(opts)
    vopt_rho = var_options(opts, array=False) if v['rho_is_scalar'] else vopt
    prec = opts.precision
    D = ('define',)
    defs = default_defines(opts)
    defs += [Row('nn_', n, True, 'uint', D), Row('mm_', m, True, 'uint', D),
             Row('nm_', n + m, True, 'uint', D), Row('dnm_', 2 * (n + m), True, 'uint', D),
             Row('nrow_AHi', v['AHi_CSR'].nrow, True, 'uint', D), Row('nrow_HiA', v['HiA_CSR'].nrow, True, 'uint', D),
             Row('NN_', N, True, 'uint', D),
             Row('k_max', int(opts.solver['k_max']), True, 'uint', D),
             Row('tol', float(opts.solver['tol']), True, prec, D)]
    if v['rho_is_scalar']:
        defs.append(Row('SCALAR_RHO', 1, False, 'bool', D))
    consts = [Row('rho', v['rho'], True, prec, vopt_rho), Row('rho_i', v['rho_i'], True, prec, vopt_rho),
              Row('Tz', v['Tz'], True, prec, vopt), Row('Sz', v['Sz'], True, prec, vopt),
              Row('LB', v['LB'], True, prec, vopt), Row('UB', v['UB'], True, prec, vopt),
              Row('L_val', v['L_CSC'].val, True, prec, vopt), Row('L_col', v['L_CSC'].col, True, 'int', vopt),
              Row('L_row', v['L_CSC'].row, True, 'int', vopt), Row('Dinv', v['Dinv'], True, prec, vopt)]
    for nm_ in ('AHi', 'HiA', 'Hi'):
        s = v[nm_ + '_CSR']
        consts += [Row(nm_ + '_val', s.val, True, prec, vopt), Row(nm_ + '_col', s.col, True, 'int', vopt),
                   Row(nm_ + '_row', s.row, True, 'int', vopt)]
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    zlen = 2 * N * (n + m)
    return SolverSpec(
        formulation='MPCT', method='ADMM', submethod='cs', func_name='MPCT_ADMM_cs', kernel='MPCT_ADMM_cs',
        defines=defs, constants=consts,
        ref_code='formulations/+MPCT/code_MPCT_ADMM_cs_C.c',
        ref_header='formulations/+MPCT/header_MPCT_ADMM_cs_C.h',
        sol_fields=(('z', zlen), ('v', zlen), ('lambda', zlen)),
        vars=v, dims=dict(n=n, m=m, N=N, dim=zlen, n_eq=v['AHi_CSR'].nrow))


# --------------------------------------------------------------------------------------
# ADMM with the semi-banded (banded + low rank, Woodbury) equality-constrained QP step (submethod 'semiband')
# --------------------------------------------------------------------------------------
def compute_MPCT_ADMM_semiband_ingredients(recipe):
    """formulations/+MPCT/compute_MPCT_ADMM_semiband_ingredients.m:20-365, scalar rho, hard constraints, no constrained output
    (the defaults of def_options_MPCT_ADMM_semiband.m).  Decision vector z = (x_0, u_0, ..., x_{N-1}, u_{N-1}, x_s, u_s); the
    Hessian is banded plus the 2 (n + m) columns / rows that couple every stage with (x_s, u_s); the z update applies the Woodbury
    identity twice (to the Hessian and to the Schur complement of the equality constraints)."""
    A, B, n, m, N = get_sys_param(recipe)
    sys, param, opt = recipe.sys, recipe.param, recipe.options
    solver = opt.solver
    nm = n + m
    if not np.isscalar(solver['rho']) or solver.get('force_vector_rho', False):
        raise NotImplementedError('MPCT_ADMM_semiband: vector rho is not generated (scalar rho only)')
    if solver.get('soft_constraints', False) or solver.get('constrained_output', False):
        raise NotImplementedError('MPCT_ADMM_semiband: soft_constraints / constrained_output are not generated')
    Q, R, T, S = (np.asarray(param[k], float) for k in ('Q', 'R', 'T', 'S'))
    inf = float(opt.inf_value)
    LBx = np.asarray(sys.get('LBx', -inf * np.ones(n)), float).ravel()
    UBx = np.asarray(sys.get('UBx', inf * np.ones(n)), float).ravel()
    LBu = np.asarray(sys.get('LBu', -inf * np.ones(m)), float).ravel()
    UBu = np.asarray(sys.get('UBu', inf * np.ones(m)), float).ravel()
    rho = float(solver['rho'])
    QR = sla.block_diag(Q, R)
    band = sla.block_diag(np.kron(np.eye(N), QR), sla.block_diag(N * Q + T, N * R + S))
    H = band.copy()
    H[:-nm, -nm:] = np.kron(np.ones((N, 1)), -QR)
    H[-nm:, :-nm] = np.kron(np.ones((1, N)), -QR)
    G = np.zeros(((N + 2) * n, (N + 1) * nm))
    G[:n, :n] = np.eye(n)                                             # x_0 = x(t)
    for l in range(N):                                                # x_{l+1} = A x_l + B u_l; the last one targets x_s
        r0, c0 = (l + 1) * n, l * nm
        G[r0:r0 + n, c0:c0 + nm + n] = np.hstack([A, B, -np.eye(n)])
    G[-n:, -nm:] = np.hstack([A - np.eye(n), B])                      # (x_s, u_s) is an equilibrium
    n_z = H.shape[0]
    Gamma_hat = band + rho * np.eye(n_z)
    Gamma_hat_inv = np.linalg.inv(Gamma_hat)
    Y = np.kron(np.ones((N, 1)), -QR)
    NNr, MM = Y.shape
    U_hat = sla.block_diag(Y, np.eye(MM))
    V_hat = np.block([[np.zeros((MM, NNr)), np.eye(MM)], [Y.T, np.zeros((MM, MM))]])
    Gamma_tilde = G @ Gamma_hat_inv @ G.T
    Gamma_tilde_inv = np.linalg.inv(Gamma_tilde)
    U_tilde_full = -G @ Gamma_hat_inv @ U_hat @ np.linalg.inv(np.eye(2 * MM) + V_hat @ Gamma_hat_inv @ U_hat)
    U_tilde = np.vstack([U_tilde_full[:2 * n, :], U_tilde_full[N * n:(N + 2) * n, :]])
    V_tilde = V_hat @ Gamma_hat_inv @ G.T
    M_hat = np.linalg.inv(np.eye(2 * nm) + V_hat @ Gamma_hat_inv @ U_hat) @ V_hat
    M_hat_x1 = np.vstack([M_hat[:n, :n], M_hat[nm:nm + n, :n]])
    M_hat_x2 = np.vstack([M_hat[:n, N * nm:N * nm + n], M_hat[nm:nm + n, N * nm:N * nm + n]])
    M_hat_u1 = np.vstack([M_hat[n:nm, n:nm], M_hat[nm + n:2 * nm, n:nm]])
    M_hat_u2 = np.vstack([M_hat[n:nm, N * nm + n:(N + 1) * nm], M_hat[nm + n:2 * nm, N * nm + n:(N + 1) * nm]])
    M_tilde_full = np.linalg.inv(np.eye(2 * nm) + V_tilde @ Gamma_tilde_inv @ U_tilde_full) @ V_tilde
    M_tilde = np.hstack([M_tilde_full[:, :2 * n], M_tilde_full[:, N * n:(N + 2) * n]])
    v = dict(n=n, m=m, N=N, rho_is_scalar=True, A=A, B=B, Q=Q, R=R, T=T, S=S, G=G, H=H,
             U_tilde=U_tilde, M_hat_x1=M_hat_x1, M_hat_x2=M_hat_x2, M_hat_u1=M_hat_u1, M_hat_u2=M_hat_u2, M_tilde=M_tilde,
             LB=np.concatenate([LBx, LBu]), UB=np.concatenate([UBx, UBu]), rho=rho, rho_i=1.0 / rho,
             Q_rho_i=np.linalg.inv(Q + rho * np.eye(n)), R_rho_i=np.linalg.inv(R + rho * np.eye(m)),
             S_rho_i=np.linalg.inv(N * R + S + rho * np.eye(m)), T_rho_i=np.linalg.inv(N * Q + T + rho * np.eye(n)))
    # blocks of the upper Cholesky factor of Gamma_tilde (N + 2 diagonal blocks with inverted diagonal, N + 1 super-diagonal ones;
    # :347-363 -- the element-wise inversion of the diagonals is what `1/vars.Beta(i,i,:)` means there)
    v['Alpha'], v['Beta'] = alpha_beta_from_chol(chol_upper(Gamma_tilde), n, N + 2)
    v.update(scaling_vars(sys, n, m))
    return v


def cons_MPCT_ADMM_semiband(recipe) -> SolverSpec:
    """formulations/+MPCT/cons_MPCT_ADMM_semiband_C.m:40-173."""
    opts = recipe.options
    v = compute_MPCT_ADMM_semiband_ingredients(recipe)
    opts = opts.copy()
    opts.force_diagonal = bool(isdiag(v['Q']) and isdiag(v['R']) and isdiag(v['T']) and isdiag(v['S']))
    n, m, N = v['n'], v['m'], v['N']
    vopt = var_options(opts)
    prec = opts.precision
    D = ('define',)
    so = opts.solver
    defs = default_defines(opts)
    defs += [Row('SOFT_CONSTRAINTS', 0, True, 'bool', D), Row('CONSTRAINED_OUTPUT', 0, True, 'bool', D),
             Row('nn_', n, True, 'uint', D), Row('mm_', m, True, 'uint', D), Row('nm_', n + m, True, 'uint', D),
             Row('NN_', N, True, 'uint', D), Row('k_max', int(so['k_max']), True, 'uint', D),
             Row('tol_p', float(so['tol_p']), True, 'float', D), Row('tol_d', float(so['tol_d']), True, 'float', D),
             Row('eps_x', float(so['epsilon_x']), True, 'float', D), Row('eps_u', float(so['epsilon_u']), True, 'float', D),
             Row('inf', float(opts.inf_value), True, 'float', D),
             Row('SCALAR_RHO', 1, False, 'bool', D), Row('rho', v['rho'], True, prec, D), Row('rho_i', v['rho_i'], True, prec, D)]
    consts = [Row(k, v[k], True, prec, vopt) for k in
              ('LB', 'UB', 'Q', 'R', 'S', 'T', 'Q_rho_i', 'R_rho_i', 'S_rho_i', 'T_rho_i', 'A', 'B', 'Alpha', 'Beta', 'U_tilde',
               'M_hat_x1', 'M_hat_x2', 'M_hat_u1', 'M_hat_u2', 'M_tilde')]
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    zlen = (N + 1) * (n + m)
    spec = SolverSpec(
        formulation='MPCT', method='ADMM', submethod='semiband', func_name='MPCT_ADMM_semiband', kernel='MPCT_ADMM_semiband',
        defines=defs, constants=consts,
        ref_code='formulations/+MPCT/code_MPCT_ADMM_semiband_C.c',
        ref_header='formulations/+MPCT/header_MPCT_ADMM_semiband_C.h',
        sol_fields=(('z', zlen), ('v', zlen), ('lambda', zlen)),
        vars=v, dims=dict(n=n, m=m, N=N, dim=zlen))
    spec.options_override = opts
    return spec
