"""laxMPC (MPC without terminal constraint) and equMPC (terminal equality) recipes.

Host-side restatement of

* formulations/+laxMPC/compute_laxMPC_FISTA_ingredients.m:22-153, cons_laxMPC_FISTA_C.m:38-135
* formulations/+laxMPC/compute_laxMPC_ADMM_ingredients.m,         cons_laxMPC_ADMM_C.m
* formulations/+equMPC/compute_equMPC_FISTA_ingredients.m,        cons_equMPC_FISTA_C.m
* formulations/+equMPC/compute_equMPC_ADMM_ingredients.m,         cons_equMPC_ADMM_C.m

The four solvers differ only by the terminal block (``T`` present for lax, the
last ``n`` columns of ``Aeq`` dropped for equ), so they share one implementation.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from .common import (Row, SolverSpec, alpha_beta_from_chol, chol_upper, default_defines,
                     dynamics_constraint, engineering_rows, get_sys_param, isdiag, scaling_vars,
                     stack_bounds, var_options)


def _bound_rows(spec_consts, defs, LB, UB, n, precision, vopt, terminal, time_varying):
    """Bounds block of the cons_* files (cons_laxMPC_FISTA_C.m:80-97,
    cons_equMPC_ADMM_C.m).  ``LB`` with more than one column means per-stage
    bounds (``VAR_BOUNDS``)."""
    if LB.ndim == 2 and LB.shape[1] > 1:
        spec_consts.append(Row('LB0', LB[n:, 0], True, precision, vopt))
        spec_consts.append(Row('UB0', UB[n:, 0], True, precision, vopt))
        if terminal:
            spec_consts.append(Row('LB', LB[:, 1:-1].T.copy(), True, precision, vopt))
            spec_consts.append(Row('UB', UB[:, 1:-1].T.copy(), True, precision, vopt))
            spec_consts.append(Row('LBN', LB[:n, -1], True, precision, vopt))
            spec_consts.append(Row('UBN', UB[:n, -1], True, precision, vopt))
        else:
            spec_consts.append(Row('LB', LB[:, 1:].T.copy(), True, precision, vopt))
            spec_consts.append(Row('UB', UB[:, 1:].T.copy(), True, precision, vopt))
        defs.append(Row('VAR_BOUNDS', 1, True, 'int', ('define',)))
    elif not time_varying:
        spec_consts.append(Row('LB', LB, True, precision, vopt))
        spec_consts.append(Row('UB', UB, True, precision, vopt))


# --------------------------------------------------------------------------------------
# FISTA
# --------------------------------------------------------------------------------------
def compute_FISTA_ingredients(recipe, terminal: bool):
    """compute_laxMPC_FISTA_ingredients.m:22-153 (``terminal=True``) /
    compute_equMPC_FISTA_ingredients.m (``terminal=False``)."""
    A, B, n, m, N = get_sys_param(recipe)
    Q = np.asarray(recipe.param['Q'], float)
    R = np.asarray(recipe.param['R'], float)
    name = 'laxMPC' if terminal else 'equMPC'
    if terminal:
        T = np.asarray(recipe.param['T'], float)
        if not isdiag(sla.block_diag(Q, R, T)):
            raise ValueError(f'Spcies:{name}:FISTA:non_diagonal: matrices Q, R and T must be diagonal')
        H = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)), T)
    else:
        if not isdiag(sla.block_diag(Q, R)):
            raise ValueError(f'Spcies:{name}:FISTA:non_diagonal: matrices Q and R must be diagonal')
        H = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)))
    Aeq = dynamics_constraint(A, B, N)
    if not terminal:
        Aeq = Aeq[:, :-n]
    v = dict(n=n, m=m, N=N)
    if not recipe.options.time_varying:
        Hinv = np.linalg.inv(H)
        W = Aeq @ Hinv @ Aeq.T
        Wc = chol_upper(W)
        v['Alpha'], v['Beta'] = alpha_beta_from_chol(Wc, n, N)
        v['AB'] = np.hstack([A, B])
        v['Q'] = -np.diag(Q)
        v['R'] = -np.diag(R)
        v['QRi'] = -np.concatenate([1.0 / np.diag(Q), 1.0 / np.diag(R)])
        v['W'] = W
    if terminal:
        v['T'] = -np.diag(T)
        v['Ti'] = -1.0 / np.diag(T)
    v['LB'], v['UB'] = stack_bounds(recipe.sys)
    v.update(scaling_vars(recipe.sys, n, m))
    v['H'] = H
    v['Aeq'] = Aeq
    return v


def _cons_FISTA(recipe, terminal: bool) -> SolverSpec:
    opts = recipe.options
    v = compute_FISTA_ingredients(recipe, terminal)
    n, m, N = v['n'], v['m'], v['N']
    name = 'laxMPC' if terminal else 'equMPC'
    if opts.time_varying and v['LB'].ndim > 1:
        raise ValueError(f'{name} FISTA time varying solver only allows fixed bounds along the prediction horizon')
    vopt = var_options(opts)
    prec = opts.precision
    defs = default_defines(opts)
    defs += [Row('nn_', n, True, 'uint', ('define',)), Row('mm_', m, True, 'uint', ('define',)),
             Row('nm_', n + m, True, 'uint', ('define',)), Row('NN_', N, True, 'uint', ('define',)),
             Row('k_max', int(opts.solver['k_max']), True, 'uint', ('define',)),
             Row('tol', float(opts.solver['tol']), True, 'float', ('define',))]
    consts = []
    _bound_rows(consts, defs, v['LB'], v['UB'], n, prec, vopt, terminal, opts.time_varying)
    if not opts.time_varying:
        for k in ('AB', 'Alpha', 'Beta', 'Q', 'R', 'QRi'):
            consts.append(Row(k, v[k], True, prec, vopt))
    if terminal:
        consts.append(Row('T', v['T'], True, prec, vopt))
        consts.append(Row('Ti', v['Ti'], True, prec, vopt))
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    zlen = N * (n + m) if terminal else N * (n + m) - n
    return SolverSpec(
        formulation=name, method='FISTA', submethod='', func_name=f'{name}_FISTA', kernel=f'{name}_FISTA',
        defines=defs, constants=consts,
        ref_code=f'formulations/+{name}/code_{name}_FISTA_C.c',
        ref_header=f'formulations/+{name}/header_{name}_FISTA_C.h',
        sol_fields=(('z', zlen), ('lambda', N * n)),
        extra_inputs=('A_in', 'B_in', 'Q_in', 'R_in', 'LB_in', 'UB_in') if opts.time_varying else (),
        vars=v, dims=dict(n=n, m=m, N=N))


def cons_laxMPC_FISTA(recipe):
    return _cons_FISTA(recipe, True)


def cons_equMPC_FISTA(recipe):
    return _cons_FISTA(recipe, False)


# --------------------------------------------------------------------------------------
# ADMM
# --------------------------------------------------------------------------------------
def compute_ADMM_ingredients(recipe, terminal: bool):
    """compute_laxMPC_ADMM_ingredients.m (``terminal=True``) /
    compute_equMPC_ADMM_ingredients.m (``terminal=False``)."""
    A, B, n, m, N = get_sys_param(recipe)
    Q = np.asarray(recipe.param['Q'], float)
    R = np.asarray(recipe.param['R'], float)
    name = 'laxMPC' if terminal else 'equMPC'
    if not isdiag(sla.block_diag(Q, R)):
        raise ValueError(f'Spcies:{name}:ADMM:non_diagonal: matrices Q and R must be diagonal')
    dimz = N * (n + m) if terminal else N * (n + m) - n
    rho = recipe.options.solver['rho']
    if np.isscalar(rho) and recipe.options.solver.get('force_vector_rho', False):
        rho = float(rho) * np.ones(dimz)
    v = dict(n=n, m=m, N=N)
    v['rho_is_scalar'] = bool(np.isscalar(rho))
    if terminal:
        T = np.asarray(recipe.param['T'], float)
        H = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)), T)
    else:
        H = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)))
    Hhat = H + (float(rho) * np.eye(dimz) if v['rho_is_scalar'] else np.diag(np.asarray(rho, float)))
    Aeq = dynamics_constraint(A, B, N)
    if not terminal:
        Aeq = Aeq[:, :-n]
    if not recipe.options.time_varying:
        Hinv = np.linalg.inv(Hhat)
        W = Aeq @ Hinv @ Aeq.T
        Wc = chol_upper(W)
        v['Alpha'], v['Beta'] = alpha_beta_from_chol(Wc, n, N)
        v['Hi_0'] = np.diag(Hinv[:m, :m]).copy()
        mid = np.diag(Hinv)[m:m + (N - 1) * (n + m)]
        v['Hi'] = mid.reshape(N - 1, n + m).copy()
        if terminal:
            v['Hi_N'] = Hinv[-n:, -n:].copy()
        v['AB'] = np.hstack([A, B])
        v['Q'] = -np.diag(Q)
        v['R'] = -np.diag(R)
        v['W'] = W
    elif terminal:
        v['T_rho_i'] = np.linalg.inv(T + float(rho) * np.eye(n))
    if terminal:
        v['T'] = -T
    v['LB'], v['UB'] = stack_bounds(recipe.sys)
    if v['rho_is_scalar']:
        v['rho'] = float(rho)
        v['rho_i'] = 1.0 / float(rho)
    else:
        rho = np.asarray(rho, float)
        v['rho_0'] = rho[:m]
        v['rho_i_0'] = 1.0 / rho[:m]
        if terminal:
            v['rho'] = rho[m:-n].reshape(N - 1, n + m)
            v['rho_i'] = (1.0 / rho[m:-n]).reshape(N - 1, n + m)
            v['rho_N'] = rho[-n:]
            v['rho_i_N'] = 1.0 / rho[-n:]
        else:
            # the reference slices rho(m+1:end-n) for equMPC too (compute_equMPC_ADMM_ingredients.m),
            # which drops the last stage; a vector rho of length N(n+m)-n has exactly (N-1)(n+m) entries
            # after the first m, so take those.
            v['rho'] = rho[m:m + (N - 1) * (n + m)].reshape(N - 1, n + m)
            v['rho_i'] = 1.0 / v['rho']
    v.update(scaling_vars(recipe.sys, n, m))
    v['H'] = H
    v['Aeq'] = Aeq
    return v


def _cons_ADMM(recipe, terminal: bool) -> SolverSpec:
    opts = recipe.options
    v = compute_ADMM_ingredients(recipe, terminal)
    n, m, N = v['n'], v['m'], v['N']
    name = 'laxMPC' if terminal else 'equMPC'
    if opts.time_varying and v['LB'].ndim > 1:
        raise ValueError(f'{name} ADMM time varying solver only allows fixed bounds along the prediction horizon')
    if opts.time_varying and not v['rho_is_scalar']:
        raise ValueError(f'{name} ADMM time varying solver only allows the use of a scalar rho')
    vopt = var_options(opts)
    prec = opts.precision
    defs = default_defines(opts)
    defs += [Row('nn_', n, True, 'uint', ('define',)), Row('mm_', m, True, 'uint', ('define',)),
             Row('nm_', n + m, True, 'uint', ('define',)), Row('NN_', N, True, 'uint', ('define',)),
             Row('k_max', int(opts.solver['k_max']), True, 'uint', ('define',)),
             Row('tol', float(opts.solver['tol']), True, 'float', ('define',))]
    consts = []
    _bound_rows(consts, defs, v['LB'], v['UB'], n, prec, vopt, terminal, opts.time_varying)
    if not opts.time_varying:
        consts.append(Row('Hi', v['Hi'], True, prec, vopt))
        consts.append(Row('Hi_0', v['Hi_0'], True, prec, vopt))
        if terminal:
            consts.append(Row('Hi_N', v['Hi_N'], True, prec, vopt))
        for k in ('Q', 'R', 'AB', 'Alpha', 'Beta'):
            consts.append(Row(k, v[k], True, prec, vopt))
    elif terminal:
        consts.append(Row('T_rho_i', v['T_rho_i'], True, prec, vopt))
    if terminal:
        consts.append(Row('T', v['T'], True, prec, vopt))
    if opts.in_engineering:
        consts += engineering_rows(v, prec, vopt)
    if v['rho_is_scalar']:
        defs.append(Row('SCALAR_RHO', 1, False, 'bool', ('define',)))
        defs.append(Row('rho', v['rho'], True, prec, ('define',)))
        defs.append(Row('rho_i', v['rho_i'], True, prec, ('define',)))
    else:
        consts.append(Row('rho', v['rho'], True, prec, vopt))
        consts.append(Row('rho_0', v['rho_0'], True, prec, vopt))
        if terminal:
            consts.append(Row('rho_N', v['rho_N'], True, prec, vopt))
        consts.append(Row('rho_i', v['rho_i'], True, prec, vopt))
        consts.append(Row('rho_i_0', v['rho_i_0'], True, prec, vopt))
        if terminal:
            consts.append(Row('rho_i_N', v['rho_i_N'], True, prec, vopt))
    zlen = N * (n + m) if terminal else N * (n + m) - n
    return SolverSpec(
        formulation=name, method='ADMM', submethod='', func_name=f'{name}_ADMM', kernel=f'{name}_ADMM',
        defines=defs, constants=consts,
        ref_code=f'formulations/+{name}/code_{name}_ADMM_C.c',
        ref_header=f'formulations/+{name}/header_{name}_ADMM_C.h',
        extra_inputs=('A_in', 'B_in', 'Q_in', 'R_in', 'LB_in', 'UB_in') if opts.time_varying else (),
        sol_fields=(('z', zlen), ('v', zlen), ('lambda', zlen)),
        vars=v, dims=dict(n=n, m=m, N=N))


def cons_laxMPC_ADMM(recipe):
    return _cons_ADMM(recipe, True)


def cons_equMPC_ADMM(recipe):
    return _cons_ADMM(recipe, False)
