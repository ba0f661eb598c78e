"""ellipMPC recipes: MPC with a terminal ellipsoidal constraint.

Host-side restatement of

* formulations/+ellipMPC/compute_ellipMPC_ADMM_soc_ingredients.m:23-215,
  cons_ellipMPC_ADMM_soc_C.m  (terminal set as a second-order cone, CSR/CSC-LDL algebra)
* formulations/+ellipMPC/compute_ellipMPC_ADMM_ingredients.m, cons_ellipMPC_ADMM_C.m
  (terminal variable in the P^(1/2) metric)
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .. import sp_utils
from .common import (Row, SolverSpec, alpha_beta_from_chol, chol_upper, default_defines,
                     dynamics_constraint, engineering_rows, get_sys_param, isdiag, scaling_vars,
                     var_options)


def _sqrtm_spd(P):
    S = sla.sqrtm(np.asarray(P, float))
    return np.real(S)                      # P is SPD; SciPy may return a complex-typed array


def _tightened_bounds(recipe, n, m, N):
    """``LB = [LBu; (LBx + incBx_i; LBu + incBu_i) for i = 2..N]``
    (compute_ellipMPC_ADMM_soc_ingredients.m:100-121)."""
    sys, param = recipe.sys, recipe.param
    incBx = np.asarray(param.get('incBx', np.zeros((n, N + 1))), float).reshape(n, N + 1)
    incBu = np.asarray(param.get('incBu', np.zeros((m, N + 1))), float).reshape(m, N + 1)
    LBx, UBx = np.asarray(sys['LBx'], float).ravel(), np.asarray(sys['UBx'], float).ravel()
    LBu, UBu = np.asarray(sys['LBu'], float).ravel(), np.asarray(sys['UBu'], float).ravel()
    LB, UB = [LBu], [UBu]
    for i in range(1, N):
        LB += [LBx + incBx[:, i], LBu + incBu[:, i]]
        UB += [UBx - incBx[:, i], UBu - incBu[:, i]]
    return np.concatenate(LB), np.concatenate(UB)


# --------------------------------------------------------------------------------------
# ADMM_soc
# --------------------------------------------------------------------------------------
def compute_ellipMPC_ADMM_soc_ingredients(recipe):
    A, B, n, m, N = get_sys_param(recipe)
    param = recipe.param
    Q, R, T, P = (np.asarray(param[k], float) for k in ('Q', 'R', 'T', 'P'))
    c = np.asarray(param.get('c', np.zeros(n)), float).ravel()
    r = float(param.get('r', 1.0))
    if not isdiag(sla.block_diag(Q, R)):
        raise ValueError('Spcies:ellipMPC:ADMM:non_diagonal: matrices Q and R must be diagonal')
    sigma, rho = float(recipe.options.solver['sigma']), float(recipe.options.solver['rho'])

    H = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)), T, 0.0)
    dim = H.shape[0]
    G = sla.block_diag(dynamics_constraint(A, B, N), 1.0)
    b = np.concatenate([np.zeros(n * N), [r]])
    n_eq = G.shape[0]
    P_half = _sqrtm_spd(P)
    qSOC = -(c @ P)
    bSOC = np.linalg.solve(P_half, -qSOC)
    C = np.hstack([np.zeros((n + 1, dim - n - 1)),
                   np.block([[np.zeros((1, n)), -np.ones((1, 1))], [-P_half, np.zeros((n, 1))]])])
    d = np.concatenate([[0.0], -bSOC])
    n_s = C.shape[0]
    LB, UB = _tightened_bounds(recipe, n, m, N)

    Hh = sla.block_diag(H + sigma * np.eye(dim), rho * np.eye(n_s))
    Gh = np.block([[G, np.zeros((n_eq, n_s))], [C, np.eye(n_s)]])
    bh = np.concatenate([b, d])
    Hhi = np.linalg.inv(Hh)
    W = Gh @ Hhi @ Gh.T
    Wc = chol_upper(W)
    Wd = np.diag(Wc)
    L = Wc.T @ np.diag(1.0 / Wd)
    Dinv = 1.0 / (Wd ** 2)
    v = dict(n=n, m=m, N=N, dim=dim, n_s=n_s, n_eq=n_eq,
             A=A, Q=-Q, R=-R, T=-T, LB=LB, UB=UB,
             rho=rho, rho_i=1.0 / rho, sigma=sigma, sigma_i=1.0 / sigma,
             L_CSC=sp_utils.full2CSC(L - np.eye(L.shape[0])), Dinv=Dinv,
             GhHhi_CSR=sp_utils.full2CSR(-Gh @ Hhi),
             HhiGh_CSR=sp_utils.full2CSR(-Hhi @ Gh.T),
             Hhi_CSR=sp_utils.full2CSR(-Hhi),
             PhiP=np.linalg.inv(P_half) @ P, P_half_i=np.linalg.inv(P_half), bh=bh,
             H=H, G=G, C=C, d=d, Hh=Hh, Gh=Gh, W=W, P=P, P_half=P_half)
    v.update(scaling_vars(recipe.sys, n, m))
    return v


def cons_ellipMPC_ADMM_soc(recipe) -> SolverSpec:
    opts = recipe.options
    v = compute_ellipMPC_ADMM_soc_ingredients(recipe)
    n, m, N = v['n'], v['m'], v['N']
    vopt = var_options(opts, array=False)          # cons_ellipMPC_ADMM_soc_C.m: {'static','constant'}
    prec = opts.precision
    D = ('define',)
    defs = default_defines(opts)
    defs += [Row('dim', v['dim'], True, 'uint', D), Row('n_s', v['n_s'], True, 'uint', D),
             Row('n_eq', v['n_eq'], True, 'uint', D),
             Row('nn_', n, True, 'uint', D), Row('mm_', m, True, 'uint', D),
             Row('nm_', n + m, True, 'uint', D), Row('NN_', N, True, 'uint', D),
             Row('nrow_GhHhi', v['GhHhi_CSR'].nrow, True, 'uint', D),
             Row('nrow_HhiGh', v['HhiGh_CSR'].nrow, True, 'uint', D),
             Row('nrow_Hhi', v['Hhi_CSR'].nrow, True, 'uint', D),
             Row('k_max', int(opts.solver['k_max']), True, 'uint', D),
             Row('tol_p', float(opts.solver['tol_p']), True, prec, D),
             Row('tol_d', float(opts.solver['tol_d']), True, prec, D)]
    consts = [Row(k, v[k], True, prec, vopt) for k in
              ('rho', 'rho_i', 'sigma', 'sigma_i', 'Q', 'R', 'T', 'A', 'LB', 'UB', 'PhiP')]
    consts += [Row('L_val', v['L_CSC'].val, True, prec, vopt),
               Row('L_col', v['L_CSC'].col, True, 'int', vopt),
               Row('L_row', v['L_CSC'].row, True, 'int', vopt),
               Row('Dinv', v['Dinv'], True, prec, vopt)]
    for nm_ in ('GhHhi', 'HhiGh', 'Hhi'):
        s = v[nm_ + '_CSR']
        consts += [Row(nm_ + '_val', s.val, True, prec, vopt),
                   Row(nm_ + '_col', s.col, True, 'int', vopt),
                   Row(nm_ + '_row', s.row, True, 'int', vopt)]
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    dim, n_s = v['dim'], v['n_s']
    return SolverSpec(
        formulation='ellipMPC', method='ADMM', submethod='soc', func_name='ellipMPC_ADMM_soc',
        kernel='ellipMPC_ADMM_soc', defines=defs, constants=consts,
        ref_code='formulations/+ellipMPC/code_ellipMPC_ADMM_soc_C.c',
        ref_header='formulations/+ellipMPC/header_ellipMPC_ADMM_soc_C.h',
        extra_inputs=('r_ellip',),
        sol_fields=(('z', dim), ('s', n_s), ('z_hat', dim), ('s_hat', n_s), ('lambda', dim), ('mu', n_s)),
        vars=v, dims=dict(n=n, m=m, N=N, dim=dim, n_s=n_s, n_eq=v['n_eq']))


# --------------------------------------------------------------------------------------
# ADMM (P-metric projection)
# --------------------------------------------------------------------------------------
def compute_ellipMPC_ADMM_ingredients(recipe):
    A, B, n, m, N = get_sys_param(recipe)
    param = recipe.param
    Q, R, T, P = (np.asarray(param[k], float) for k in ('Q', 'R', 'T', 'P'))
    c = np.asarray(param.get('c', np.zeros(n)), float).ravel()
    r = float(param.get('r', 1.0))
    if not isdiag(sla.block_diag(Q, R)):
        raise ValueError('Spcies:ellipMPC:ADMM:non_diagonal: matrices Q and R must be diagonal')
    rho = recipe.options.solver['rho']
    if np.isscalar(rho) and recipe.options.solver.get('force_vector_rho', False):
        # the reference reads an undefined ``options.rho`` here; the intent is clear
        rho = float(rho) * np.ones(N * (n + m))
    scalar = bool(np.isscalar(rho))
    Hz = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)), T)
    P_half = _sqrtm_spd(P)
    scale = sla.block_diag(np.eye(Hz.shape[0] - n), P)
    # MATLAB ``rho.*M`` with a column vector rho scales rows
    H = Hz + (float(rho) * scale if scalar else np.asarray(rho, float)[:, None] * scale)
    Aeq = dynamics_constraint(A, B, N)
    Hinv = np.linalg.inv(H)
    W = Aeq @ Hinv @ Aeq.T
    Wc = chol_upper(W)
    LBz, UBz = _tightened_bounds(recipe, n, m, N)
    v = dict(n=n, m=m, N=N, rho_is_scalar=scalar)
    v['Alpha'], v['Beta'] = alpha_beta_from_chol(Wc, n, N)
    v['Hi_0'] = np.diag(Hinv[:m, :m]).copy()
    v['Hi'] = np.diag(Hinv)[m:m + (N - 1) * (n + m)].reshape(N - 1, n + m).copy()
    v['Hi_N'] = Hinv[-n:, -n:].copy()
    v['AB'] = np.hstack([A, B])
    v['LBu0'], v['UBu0'] = LBz[:m], UBz[:m]
    v['LBz'] = LBz[m:].reshape(N - 1, n + m)
    v['UBz'] = UBz[m:].reshape(N - 1, n + m)
    v['P'] = P
    v['P_half'] = P_half
    v['Pinv_half'] = np.linalg.inv(P) @ P_half
    v['Q'] = -np.diag(Q)
    v['R'] = -np.diag(R)
    v['T'] = -T
    v['c'] = c
    v['r'] = r
    if scalar:
        v['rho'] = float(rho)
        v['rho_i'] = 1.0 / float(rho)
    else:
        rho = np.asarray(rho, float)
        v['rho_0'], v['rho_N'] = rho[:m], rho[-n:]
        v['rho'] = rho[m:-n].reshape(N - 1, n + m)
        v['rho_i_0'], v['rho_i_N'] = 1.0 / rho[:m], 1.0 / rho[-n:]
        v['rho_i'] = 1.0 / v['rho']
    v.update(scaling_vars(recipe.sys, n, m))
    v['H'], v['Aeq'], v['W'] = H, Aeq, W
    return v


def cons_ellipMPC_ADMM(recipe) -> SolverSpec:
    opts = recipe.options
    v = compute_ellipMPC_ADMM_ingredients(recipe)
    n, m, N = v['n'], v['m'], v['N']
    vopt = var_options(opts)
    prec = opts.precision
    D = ('define',)
    defs = default_defines(opts)
    defs += [Row('nn_', n, True, 'uint', D), Row('mm_', m, True, 'uint', D),
             Row('nm_', n + m, True, 'uint', D), Row('NN_', N, True, 'uint', D),
             Row('k_max', int(opts.solver['k_max']), True, 'uint', D),
             Row('tol', float(opts.solver['tol']), True, 'float', D)]
    consts = [Row(k, v[k], True, prec, vopt) for k in
              ('LBu0', 'UBu0', 'LBz', 'UBz', 'Hi', 'Hi_0', 'Hi_N', 'AB', 'P', 'P_half', 'Pinv_half',
               'Alpha', 'Beta', 'Q', 'R', 'T')]
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    variables = [Row('c', v['c'], True, prec, ('variable',)), Row('r', v['r'], True, prec, ('variable',))]
    if v['rho_is_scalar']:
        defs += [Row('SCALAR_RHO', 1, False, 'bool', D), Row('rho', v['rho'], True, prec, D),
                 Row('rho_i', v['rho_i'], True, prec, D)]
    else:
        consts += [Row(k, v[k], True, prec, vopt) for k in ('rho', 'rho_0', 'rho_N', 'rho_i', 'rho_i_0', 'rho_i_N')]
    zlen = N * (n + m)
    return SolverSpec(
        formulation='ellipMPC', method='ADMM', submethod='', func_name='ellipMPC_ADMM', kernel='ellipMPC_ADMM',
        defines=defs, constants=consts, variables=variables,
        ref_code='formulations/+ellipMPC/code_ellipMPC_ADMM_C.c',
        ref_header='formulations/+ellipMPC/header_ellipMPC_ADMM_C.h',
        sol_fields=(('z', zlen), ('v', zlen), ('lambda', zlen)),
        vars=v, dims=dict(n=n, m=m, N=N))
