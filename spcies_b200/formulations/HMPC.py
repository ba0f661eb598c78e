"""HMPC (harmonic MPC) -- ADMM_split / SADMM_split recipes.

Host-side restatement of formulations/+HMPC/compute_HMPC_ADMM_split_ingredients.m:23-325
and cons_HMPC_ADMM_split_C.m:43-162 (cons_HMPC_SADMM_split_C.m:45 delegates to it; SADMM
only adds the ``alpha_SADMM`` / ``IS_SYMMETRIC`` defines).

Known reference defects handled here:
* cons_HMPC_ADMM_split_C.m:95 reads a non-existent ``recipe.solver_options``; the intent
  (``recipe.options.solver.box_constraints``) is used.
* the sparse (``sparse=true``) path relies on MATLAB ``ldl`` with pivoting; only the
  default dense path (``NON_SPARSE``: ``M1``, ``M2``) is generated.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .common import (Row, SolverSpec, default_defines, dynamics_constraint, engineering_rows,
                     get_sys_param, scaling_vars, var_options)


def compute_HMPC_ADMM_split_ingredients(recipe, box_constraints=True):
    A, B, n, m, N = get_sys_param(recipe)
    sys, param, solver = recipe.sys, recipe.param, recipe.options.solver
    nm = n + m
    LBx, UBx = np.asarray(sys['LBx'], float).ravel(), np.asarray(sys['UBx'], float).ravel()
    LBu, UBu = np.asarray(sys['LBu'], float).ravel(), np.asarray(sys['UBu'], float).ravel()
    if not box_constraints:
        E, F = np.asarray(sys['E'], float), np.asarray(sys['F'], float)
        LBy, UBy = np.asarray(sys['LBy'], float).ravel(), np.asarray(sys['UBy'], float).ravel()
    else:
        E = np.vstack([np.eye(n), np.zeros((m, n))])
        F = np.vstack([np.zeros((n, m)), np.eye(m)])
        LBy, UBy = np.concatenate([LBx, LBu]), np.concatenate([UBx, UBu])
    n_y = LBy.size
    w = float(param['w'])
    Q, R, Te, Th, Se, Sh = (np.asarray(param[k], float) for k in ('Q', 'R', 'Te', 'Th', 'Se', 'Sh'))
    use_soc = bool(solver.get('use_soc', False))

    js = np.arange(N)
    s_j, c_j = np.sin(w * js), np.cos(w * js)
    # accumulate in the reference's order (:84-95) so the sums match to the last bit
    s_sum = c_sum = s2_sum = c2_sum = sc_sum = 0.0
    for j in range(N):
        s_sum += np.sin(w * j)
        c_sum += np.cos(w * j)
        s2_sum += np.sin(w * j) ** 2
        c2_sum += np.cos(w * j) ** 2
        sc_sum += np.sin(w * j) * np.cos(w * j)

    H11 = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)))
    H12 = np.zeros(((N - 1) * nm + m, 3 * n))
    for j in range(N - 1):
        H12[j * nm + m:(j + 1) * nm, :] = np.kron([[1.0, s_j[j + 1], c_j[j + 1]]], -Q)
    H13 = np.zeros(((N - 1) * nm + m, 3 * m))
    for j in range(N):
        H13[j * nm:j * nm + m, :] = np.kron([[1.0, s_j[j], c_j[j]]], -R)
    H22 = np.block([[Te + N * Q, s_sum * Q, c_sum * Q],
                    [s_sum * Q, Th + s2_sum * Q, sc_sum * Q],
                    [c_sum * Q, sc_sum * Q, Th + c2_sum * Q]])
    H33 = np.block([[Se + N * R, s_sum * R, c_sum * R],
                    [s_sum * R, Sh + s2_sum * R, sc_sum * R],
                    [c_sum * R, sc_sum * R, Sh + c2_sum * R]])
    H23 = np.zeros((3 * n, 3 * m))
    H = np.block([[H11, H12, H13], [H12.T, H22, H23], [H13.T, H23.T, H33]])
    dim = H.shape[0]

    G = dynamics_constraint(A, B, N, n_identities=N - 2)
    In = np.eye(n)
    G = np.hstack([G, np.vstack([np.zeros((G.shape[0] - n, 3 * nm)),
                                 np.hstack([-In, -In * np.sin(w * N), -In * np.cos(w * N), np.zeros((n, 3 * m))])])])
    Zn, Znm = np.zeros((n, n)), np.zeros((n, m))
    harm = np.block([[A - In, Zn, Zn, B, Znm, Znm],
                     [Zn, A - np.cos(w) * In, np.sin(w) * In, Znm, B, Znm],
                     [Zn, -np.sin(w) * In, A - np.cos(w) * In, Znm, Znm, B]])
    G = np.vstack([G, np.hstack([np.zeros((3 * n, G.shape[1] - 3 * nm)), harm])])
    n_eq = G.shape[0]
    b = np.zeros(n * (N + 3))

    def soc_blocks():
        C_aux, dsoc = [], []
        for j in range(n_y):
            Ej, Fj = E[j:j + 1, :], F[j:j + 1, :]
            Eub, Elb = sla.block_diag(Ej, -Ej, -Ej), sla.block_diag(-Ej, -Ej, -Ej)
            Fub, Flb = sla.block_diag(Fj, -Fj, -Fj), sla.block_diag(-Fj, -Fj, -Fj)
            C_aux.append(np.vstack([np.hstack([Eub, Fub]), np.hstack([Elb, Flb])]))
            dsoc += [UBy[j], 0.0, 0.0, -LBy[j], 0.0, 0.0]
        return np.vstack(C_aux), np.array(dsoc), 2 * n_y

    if not box_constraints:
        if use_soc:
            C_aux, dsoc, n_soc = soc_blocks()
        else:
            C_aux = np.vstack([np.hstack([np.kron(np.eye(3), -E[j:j + 1, :]), np.kron(np.eye(3), -F[j:j + 1, :])])
                               for j in range(n_y)])
            dsoc, n_soc = np.zeros(3 * n_y), n_y
        C = sla.block_diag(-F, np.kron(np.eye(N - 1), np.hstack([-E, -F])), C_aux)
        d = np.concatenate([np.zeros(N * n_y), dsoc])
    else:
        if use_soc:
            C_aux, dsoc, n_soc = soc_blocks()
        else:
            C_n = np.vstack([np.kron(np.eye(3), -np.eye(n)[j:j + 1, :]) for j in range(n)])
            C_m = np.vstack([np.kron(np.eye(3), -np.eye(m)[j:j + 1, :]) for j in range(m)])
            C_aux = sla.block_diag(C_n, C_m)
            dsoc, n_soc = np.zeros(3 * n_y), n_y
        C = np.hstack([np.zeros((3 * n_soc, dim - 3 * nm)), C_aux])
        d = dsoc
    n_s = C.shape[0]

    sigma, rho = float(solver['sigma']) or 1.0, float(solver['rho'])      # (ellipHMPC reuses the blocks with its sigma = 0)
    Hh = sla.block_diag(H + sigma * np.eye(dim), rho * np.eye(n_s))
    Gh = np.block([[G, np.zeros((n_eq, n_s))], [C, np.eye(n_s)]])
    bh = np.concatenate([b, d])
    Hhi = np.linalg.inv(Hh)
    W = Gh @ Hhi @ Gh.T
    Wi = np.linalg.inv(W)
    M1 = Hhi @ Gh.T @ Wi @ Gh @ Hhi - Hhi
    M2 = Hhi @ Gh.T @ Wi

    v = dict(dim=dim, n_s=n_s, n_eq=n_eq, N=N, n=n, m=m, n_y=n_y, n_soc=n_soc,
             Q=Q, Te=Te, Se=Se, A=A,
             LB=np.concatenate([LBu, np.kron(np.ones(N - 1), np.concatenate([LBx, LBu]))]),
             UB=np.concatenate([UBu, np.kron(np.ones(N - 1), np.concatenate([UBx, UBu]))]),
             LBy=LBy, UBy=UBy, E=E, F=F, H=H, Hh=Hh, G=G, b=b, C=C, d=d, Gh=Gh, bh=bh,
             M1=M1, M2=M2, rho=rho, rho_i=1.0 / rho, sigma=sigma, sigma_i=1.0 / sigma,
             k_max=int(solver['k_max']), tol_p=float(solver['tol_p']), tol_d=float(solver['tol_d']),
             use_soc=use_soc, box_constraints=box_constraints, w=w)
    v.update(scaling_vars(sys, n, m))
    return v


def _kkt_ldl(v):
    """``sparse = true`` (compute_HMPC_ADMM_split_ingredients.m:226-236, :277-286): L D L' of the KKT matrix
    ``M = [Hh, Gh'; Gh, 0]`` for the QDLDL-style solve of code_HMPC_ADMM_split_C.c:193-209.  The reference calls MATLAB's pivoted
    ``ldl`` and *assumes* what it needs from the result (D diagonal, no permutation of the primal block, :245-248).  M is
    quasi-definite (Hh > 0, Gh of full row rank), so the factorisation without pivoting exists with D diagonal -- positive on the
    primal block, negative on the multiplier block -- and is what is computed here: no permutation at all, ``idx_x0 = 0..n-1``
    and ``bh`` in its natural order.  The instantiated C template solves the same system with the same L, Dinv."""
    from .. import sp_utils
    Hh, Gh = v['Hh'], v['Gh']
    nh, ng = Hh.shape[0], Gh.shape[0]
    M = np.block([[Hh, Gh.T], [Gh, np.zeros((ng, ng))]])
    nM = M.shape[0]
    L = np.eye(nM)
    Dg = np.zeros(nM)
    for j in range(nM):
        w = L[j, :j] * Dg[:j]
        Dg[j] = M[j, j] - L[j, :j] @ w
        if j + 1 < nM:
            L[j + 1:, j] = (M[j + 1:, j] - L[j + 1:, :j] @ w) / Dg[j]
    if not (np.all(Dg[:nh] > 0) and np.all(Dg[nh:] < 0)):
        raise ValueError('HMPC sparse: the KKT matrix is not quasi-definite')
    err = np.max(np.abs(L @ np.diag(Dg) @ L.T - M))
    if err > 1e-8 * max(1.0, np.max(np.abs(M))):
        raise ValueError('HMPC sparse: L D L\' factorisation failed (%.2e)' % err)
    Lm = L - np.eye(nM)
    Lm[np.abs(Lm) < 1e-300] = 0.0
    return dict(L_CSC=sp_utils.full2CSC(Lm), Dinv=1.0 / Dg, idx_x0=np.arange(v['n'], dtype=np.int32), KKT=M)


def _cons_split(recipe, symmetric: bool) -> SolverSpec:
    opts = recipe.options
    solver = opts.solver
    box = solver.get('box_constraints', None)
    if box is None or (isinstance(box, (list, tuple)) and len(box) == 0):
        box = 'E' not in recipe.sys                     # cons_HMPC_ADMM_split_C.m:56-62
    box = bool(box)
    sparse = bool(solver.get('sparse', False))
    v = compute_HMPC_ADMM_split_ingredients(recipe, box)
    if sparse:
        v.update(_kkt_ldl(v))
    n, m, N, dim, n_s, n_eq = v['n'], v['m'], v['N'], v['dim'], v['n_s'], v['n_eq']
    vopt = var_options(opts)
    vopt_pen = var_options(opts, array=False)
    prec = opts.precision
    D = ('define',)
    defs = default_defines(opts)
    defs += [Row('nn_', n, True, 'uint', D), Row('mm_', m, True, 'uint', D), Row('nm_', n + m, True, 'uint', D)]
    if not box:
        defs.append(Row('n_y', v['n_y'], True, 'uint', D))
    defs += [Row('NN_', N, True, 'uint', D), Row('dim', dim, True, 'uint', D), Row('n_s', n_s, True, 'uint', D),
             Row('n_eq', n_eq, True, 'uint', D), Row('n_soc', v['n_soc'], True, 'uint', D),
             Row('nrow_M', dim + n_eq + 2 * n_s, True, 'uint', D) if sparse else Row('NON_SPARSE', 1, True, 'bool', D)]
    if not box:
        defs.append(Row('COUPLED_CONSTRAINTS', 1, True, 'bool', D))
    defs += [Row('k_max', int(solver['k_max']), True, 'uint', D),
             Row('tol_p', float(solver['tol_p']), True, prec, D),
             Row('tol_d', float(solver['tol_d']), True, prec, D)]
    if symmetric:
        defs += [Row('alpha_SADMM', float(solver['alpha']), True, prec, D),
                 Row('IS_SYMMETRIC', 1, True, prec, D)]
    if v['use_soc']:
        defs.append(Row('USE_SOC', 1, True, prec, D))
    consts = [Row(k, v[k], True, prec, vopt_pen) for k in ('rho', 'rho_i', 'sigma', 'sigma_i')]
    consts += [Row('A', v['A'], True, prec, vopt), Row('QQ', v['Q'], True, prec, vopt),
               Row('Te', v['Te'], True, prec, vopt), Row('Se', v['Se'], True, prec, vopt),
               Row('LB', v['LB'], True, prec, vopt), Row('UB', v['UB'], True, prec, vopt),
               Row('LBy', v['LBy'], True, prec, vopt), Row('UBy', v['UBy'], True, prec, vopt)]
    if sparse:                                          # cons_HMPC_ADMM_split_C.m:136-142
        consts += [Row('L_val', v['L_CSC'].val, True, prec, vopt), Row('L_col', v['L_CSC'].col, True, 'int', vopt),
                   Row('L_row', v['L_CSC'].row, True, 'int', vopt), Row('Dinv', v['Dinv'], True, prec, vopt),
                   Row('idx_x0', v['idx_x0'], True, 'int', vopt)]
    else:
        consts.append(Row('M1', v['M1'], True, prec, vopt))
    if sparse:
        pass
    elif v['use_soc']:
        consts.append(Row('M2', v['M2'], True, prec, vopt))
        defs.append(Row('dim_M2', n_eq + n_s, True, prec, D))
    else:
        consts.append(Row('M2', v['M2'][:, :n].copy(), True, prec, vopt))
        defs.append(Row('dim_M2', n, True, prec, D))
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    variables = [Row('bh', v['bh'], True, prec, ())]
    method = 'SADMM' if symmetric else 'ADMM'
    return SolverSpec(
        formulation='HMPC', method=method, submethod='split', func_name='HMPC_ADMM', kernel='HMPC_ADMM_split',
        defines=defs, constants=consts, variables=variables,
        ref_code='formulations/+HMPC/code_HMPC_ADMM_split_C.c',
        ref_header='formulations/+HMPC/header_HMPC_ADMM_split_C.h',
        sol_fields=(('z', dim), ('s', n_s), ('z_hat', dim), ('s_hat', n_s), ('lambda', dim), ('mu', n_s)),
        vars=v, dims=dict(n=n, m=m, N=N, dim=dim, n_s=n_s, n_eq=n_eq))


def cons_HMPC_ADMM_split(recipe):
    return _cons_split(recipe, False)


def cons_HMPC_SADMM_split(recipe):
    return _cons_split(recipe, True)


# --------------------------------------------------------------------------------------
# ADMM without the (z_hat, s_hat) splitting: HMPC / ADMM / '' (the toolbox default for HMPC,
# classes/Spcies_options.m:94,104) and ellipHMPC / ADMM
# --------------------------------------------------------------------------------------
def compute_HMPC_ADMM_ingredients(recipe, box_constraints=True, ellip=False):
    """formulations/+HMPC/compute_HMPC_ADMM_ingredients.m:25-332 and compute_ellipHMPC_ADMM_ingredients.m:25-288.

    Same Hessian / equality blocks as the split solver; the constraints enter through ``s = -(C z [- d])`` with
    ``C = blkdiag(box part, cone part)``, ``Hh = H + rho C'C``, ``M1 = Hh^-1 G' W^-1 G Hh^-1 - Hh^-1``,
    ``M2 = Hh^-1 G' W^-1`` (first n columns).  MATLAB's pivoted ``ldl`` of the KKT matrix (:246-253) only feeds the
    ``sparse`` variant, which the C template of this solver does not have."""
    from .. import sp_utils
    base = compute_HMPC_ADMM_split_ingredients(recipe, box_constraints and not ellip)
    A, B, n, m, N = get_sys_param(recipe)
    sys, param, solver = recipe.sys, recipe.param, recipe.options.solver
    nm = n + m
    H, G, dim, n_eq = base['H'], base['G'], base['dim'], base['n_eq']
    E, F, LBy, UBy = base['E'], base['F'], base['LBy'], base['UBy']
    n_y = LBy.size
    use_soc = bool(solver.get('use_soc', False))
    box = box_constraints and not ellip

    if use_soc:
        C_aux, dsoc = [], []
        for j in range(n_y):
            Ej, Fj = E[j:j + 1, :], F[j:j + 1, :]
            Eub, Elb = sla.block_diag(Ej, -Ej, -Ej), sla.block_diag(-Ej, -Ej, -Ej)
            Fub, Flb = sla.block_diag(Fj, -Fj, -Fj), sla.block_diag(-Fj, -Fj, -Fj)
            C_aux.append(np.vstack([np.hstack([Eub, Fub]), np.hstack([Elb, Flb])]))
            dsoc += [UBy[j], 0.0, 0.0, -LBy[j], 0.0, 0.0]
        C_aux, dsoc, n_soc = np.vstack(C_aux), np.array(dsoc), 2 * n_y
    elif box:
        C_n = np.vstack([np.kron(np.eye(3), -np.eye(n)[j:j + 1, :]) for j in range(n)])
        C_m = np.vstack([np.kron(np.eye(3), -np.eye(m)[j:j + 1, :]) for j in range(m)])
        C_aux, dsoc, n_soc = sla.block_diag(C_n, C_m), np.zeros(3 * n_y), n_y
    else:
        C_aux = np.vstack([np.hstack([np.kron(np.eye(3), -E[j:j + 1, :]), np.kron(np.eye(3), -F[j:j + 1, :])])
                           for j in range(n_y)])
        dsoc, n_soc = np.zeros(3 * n_y), n_y
    if box:
        C = sla.block_diag(-np.eye(m), np.kron(np.eye(N - 1), sla.block_diag(-np.eye(n), -np.eye(m))), C_aux)
        d = np.concatenate([np.zeros((N - 1) * nm + m), dsoc])
        LB, UB = base['LB'], base['UB']
        n_box = (N - 1) * nm + m
    else:
        C = sla.block_diag(-F, np.kron(np.eye(N - 1), np.hstack([-E, -F])), C_aux)
        d = np.concatenate([np.zeros(N * n_y), dsoc])
        LB, UB = np.kron(np.ones(N), LBy), np.kron(np.ones(N), UBy)
        n_box = N * n_y
    n_s = C.shape[0]
    rho = float(solver['rho'])
    Hh = H + rho * (C.T @ C)
    Hhi = np.linalg.inv(Hh)
    W = G @ Hhi @ G.T
    Wi = np.linalg.inv(W)
    M1 = Hhi @ G.T @ Wi @ G @ Hhi - Hhi
    M2 = (Hhi @ G.T @ Wi)[:, :n].copy()
    v = dict(base)
    v.update(dim=dim, n_s=n_s, n_eq=n_eq, n_y=n_y, n_soc=n_soc, n_box=n_box, C=C, d=d, LB=LB, UB=UB, Hh=Hh, M1=M1, M2=M2,
             C_CSR=sp_utils.full2CSR(C), Ct_CSR=sp_utils.full2CSR(C.T), rho=rho, rho_i=1.0 / rho, use_soc=use_soc,
             box_constraints=box, R=np.asarray(param['R'], float), Th=np.asarray(param['Th'], float),
             Sh=np.asarray(param['Sh'], float))
    if ellip:                                   # compute_ellipHMPC_ADMM_ingredients.m:222-223 (sic: tightened by sigma)
        sg = float(solver.get('sigma', 0.0))
        v['LBy'], v['UBy'] = LBy + sg, UBy - sg
    return v


def _cons_nonsplit(recipe, ellip: bool) -> SolverSpec:
    opts = recipe.options
    solver = opts.solver
    box = solver.get('box_constraints', None)
    if box is None or (isinstance(box, (list, tuple)) and len(box) == 0):
        box = 'E' not in recipe.sys                     # cons_HMPC_ADMM_C.m:57-63
    box = bool(box) and not ellip
    v = compute_HMPC_ADMM_ingredients(recipe, box, ellip)
    n, m, N, dim, n_s, n_eq = v['n'], v['m'], v['N'], v['dim'], v['n_s'], v['n_eq']
    vopt = var_options(opts)
    vopt_pen = var_options(opts, array=False)
    prec = opts.precision
    D = ('define',)
    defs = default_defines(opts)
    defs += [Row('nn_', n, True, 'uint', D), Row('mm_', m, True, 'uint', D), Row('nm_', n + m, True, 'uint', D),
             Row('NN_', N, True, 'uint', D), Row('dim', dim, True, 'uint', D), Row('n_s', n_s, True, 'uint', D),
             Row('n_eq', n_eq, True, 'uint', D), Row('n_soc', v['n_soc'], True, 'uint', D),
             Row('n_y', v['n_y'], True, 'uint', D), Row('n_box', v['n_box'], True, 'uint', D),
             Row('nrow_C', v['C_CSR'].nrow, True, 'uint', D), Row('nrow_Ct', v['Ct_CSR'].nrow, True, 'uint', D),
             Row('k_max', int(solver['k_max']), True, 'uint', D),
             Row('tol_p', float(solver['tol_p']), True, prec, D), Row('tol_d', float(solver['tol_d']), True, prec, D)]
    if opts.method == 'SADMM':
        defs += [Row('alpha_SADMM', float(solver['alpha']), True, prec, D), Row('IS_SYMMETRIC', 1, True, prec, D)]
    if v['use_soc']:
        defs.append(Row('USE_SOC', 1, True, prec, D))
    consts = [Row('rho', v['rho'], True, prec, vopt_pen), Row('rho_i', v['rho_i'], True, prec, vopt_pen),
              Row('A', v['A'], True, prec, vopt)]
    for nm_ in ('C', 'Ct'):
        s = v[nm_ + '_CSR']
        consts += [Row(nm_ + '_val', s.val, True, prec, vopt), Row(nm_ + '_col', s.col, True, 'int', vopt),
                   Row(nm_ + '_row', s.row, True, 'int', vopt)]
    consts += [Row('QQ', v['Q'], True, prec, vopt), Row('Te', v['Te'], True, prec, vopt)]
    if ellip:
        consts.append(Row('Th', v['Th'], True, prec, vopt))
    consts.append(Row('Se', v['Se'], True, prec, vopt))
    if ellip:
        consts.append(Row('Sh', v['Sh'], True, prec, vopt))
    consts += [Row('LB', v['LB'], True, prec, vopt), Row('UB', v['UB'], True, prec, vopt),
               Row('LBy', v['LBy'], True, prec, vopt), Row('UBy', v['UBy'], True, prec, vopt)]
    if v['use_soc']:
        consts.append(Row('d', v['d'], True, prec, vopt))
    consts += [Row('M1', v['M1'], True, prec, vopt), Row('M2', v['M2'], True, prec, vopt)]
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    if ellip:
        return SolverSpec(
            formulation='ellipHMPC', method='ADMM', submethod='', func_name='HMPC_ADMM', kernel='ellipHMPC_ADMM',
            defines=defs, constants=consts,
            ref_code='formulations/+HMPC/code_ellipHMPC_ADMM_C.c',
            ref_header='formulations/+HMPC/header_ellipHMPC_ADMM_C.h',
            extra_inputs=('xrs', 'xrc', 'urs', 'urc'),
            sol_fields=(('z', dim), ('s', n_s), ('lambda', n_s)),
            vars=v, dims=dict(n=n, m=m, N=N, dim=dim, n_s=n_s, n_eq=n_eq))
    return SolverSpec(
        formulation='HMPC', method=opts.method, submethod='', func_name='HMPC_ADMM', kernel='HMPC_ADMM',
        defines=defs, constants=consts,
        ref_code='formulations/+HMPC/code_HMPC_ADMM_C.c',
        ref_header='formulations/+HMPC/header_HMPC_ADMM_C.h',
        sol_fields=(('z', dim), ('s', n_s), ('lambda', n_s)),
        vars=v, dims=dict(n=n, m=m, N=N, dim=dim, n_s=n_s, n_eq=n_eq))


def cons_HMPC_ADMM(recipe):
    return _cons_nonsplit(recipe, False)


def cons_ellipHMPC_ADMM(recipe):
    return _cons_nonsplit(recipe, True)
