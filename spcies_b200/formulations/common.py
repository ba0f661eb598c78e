"""Shared pieces of the per-formulation recipes.

``Row`` mirrors one row of the reference's variable tables
(``{name, value, initialize, type, options}``, +sp_utils/add_line.m:15-24,
platforms/+C_code/dec_var.m:4-13).  ``SolverSpec`` is the platform-neutral part of
what a ``cons_<F>_<method>_<platform>.m`` builds: the ``#define`` table, the
constant table, the (non-const) variable table and the ingredients struct.
The CUDA platform (``platforms/cuda_code.py``) and the test oracle both consume
it, so both sides see the *same* numbers.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np
import scipy.linalg as sla


@dataclass
class Row:
    name: str
    value: Any
    initialize: bool = True
    type: str = 'double'
    options: tuple = ()


@dataclass
class SolverSpec:
    formulation: str
    method: str
    submethod: str
    func_name: str                 # C symbol of the single-instance solver, e.g. 'laxMPC_FISTA'
    kernel: str                    # kernel template under csrc/, e.g. 'laxMPC_FISTA'
    defines: list = field(default_factory=list)
    constants: list = field(default_factory=list)
    variables: list = field(default_factory=list)
    ref_code: str = ''             # reference template (relative to the reference root)
    ref_header: str = ''
    extra_inputs: tuple = ()       # run-time inputs after ur_in, e.g. ('r_ellip',)
    sol_fields: tuple = ()         # (name, length) of the sol_<name> debug payload
    vars: dict = field(default_factory=dict)
    dims: dict = field(default_factory=dict)
    model: tuple = ()              # (A, B) of the prediction model (the default plant of the closed-loop entry point)

    def define(self, name, default=None):
        for r in self.defines:
            if r.name == name:
                return r.value
        return default

    def const(self, name):
        for r in self.constants + self.variables:
            if r.name == name:
                return r.value
        raise KeyError(name)


def var_options(options, array=True):
    """``{'static','constant','array'}`` selection of the cons_* files
    (e.g. cons_laxMPC_FISTA_C.m:58-63)."""
    out = []
    if options.const_are_static:
        out.append('static')
    out.append('constant')
    if array:
        out.append('array')
    return tuple(out)


def default_defines(options):
    return [Row(n, v, init, t, ('define',)) for (n, v, init, t) in options.default_defCell()]


def get_sys_param(recipe):
    """Unpack ``controller.sys`` / ``controller.param`` the way every
    compute_*_ingredients.m does for plain structs (e.g.
    compute_laxMPC_FISTA_ingredients.m:35-47)."""
    sys, param = recipe.sys, recipe.param
    A = np.asarray(sys['A'], dtype=float)
    B = np.asarray(sys['Bu'] if 'Bu' in sys else sys['B'], dtype=float)
    n, m = A.shape[0], B.shape[1]
    return A, B, n, m, int(param['N'])


def isdiag(M):
    M = np.asarray(M)
    return np.count_nonzero(M - np.diag(np.diagonal(M))) == 0


def dynamics_constraint(A, B, N, n_identities=None):
    """Equality-constraint matrix of the prediction model.

    Restates the construction of e.g. compute_laxMPC_FISTA_ingredients.m:62-68:
    ``kron(eye(N-1), [A B])``, ``-I`` blocks written one block-column to the
    right (MATLAB silently *grows* the matrix by ``n`` columns on the last write),
    then the initial-condition block row ``[B, -I, 0...]`` on top.  Result is
    ``N n  x  N (n+m)`` for decision vector ``(u_0, x_1, u_1, ..., x_{N-1}, u_{N-1}, x_N)``.

    ``n_identities`` limits how many ``-I`` blocks are written (HMPC writes only
    ``N-2`` and therefore does not grow, compute_HMPC_ADMM_split_ingredients.m:132-137).
    """
    n, m = B.shape
    nm = n + m
    n_id = N - 1 if n_identities is None else n_identities
    width = (N - 1) * nm + (n if n_id == N - 1 else 0)
    G = np.zeros(((N - 1) * n, width))
    AB = np.hstack([A, B])
    for j in range(N - 1):
        G[j * n:(j + 1) * n, j * nm:(j + 1) * nm] = AB
    for j in range(n_id):
        c0 = j * nm + nm
        G[j * n:(j + 1) * n, c0:c0 + n] = -np.eye(n)
    top = np.hstack([B, -np.eye(n), np.zeros((n, G.shape[1] - n))])
    return np.vstack([top, np.hstack([np.zeros((G.shape[0], m)), G])])


def alpha_beta_from_chol(Wc, n, N):
    """Blocks of the upper Cholesky factor of the block-tridiagonal ``W``:
    ``Beta[:,:,i]`` diagonal blocks with **inverted diagonal**, ``Alpha[:,:,i]``
    super-diagonal blocks (compute_laxMPC_FISTA_ingredients.m:137-150).
    Returned as ``[block][row][col]`` arrays -- the layout dec_var.m:115-116
    emits for a MATLAB ``(row, col, block)`` array."""
    Beta = np.zeros((N, n, n))
    Alpha = np.zeros((N - 1, n, n))
    for i in range(N):
        blk = Wc[i * n:(i + 1) * n, i * n:(i + 1) * n].copy()
        for j in range(n):
            blk[j, j] = 1.0 / blk[j, j]
        Beta[i] = blk
    for i in range(N - 1):
        Alpha[i] = Wc[i * n:(i + 1) * n, (i + 1) * n:(i + 2) * n]
    return Alpha, Beta


def chol_upper(M):
    return sla.cholesky(M, lower=False)


def scaling_vars(sys, n, m):
    """Scaling vectors / operating point block shared by all ingredients files
    (e.g. compute_laxMPC_FISTA_ingredients.m:105-131)."""
    Nx = np.asarray(sys.get('Nx', np.ones(n)), dtype=float).ravel()
    Nu = np.asarray(sys.get('Nu', np.ones(m)), dtype=float).ravel()
    return dict(scaling_x=Nx, scaling_u=Nu, scaling_i_u=1.0 / Nu,
                OpPoint_x=np.asarray(sys.get('x0', np.zeros(n)), dtype=float).ravel(),
                OpPoint_u=np.asarray(sys.get('u0', np.zeros(m)), dtype=float).ravel())


def engineering_rows(vars_, precision, vopt):
    return [Row(k, vars_[k], True, precision, vopt)
            for k in ('scaling_x', 'scaling_u', 'scaling_i_u', 'OpPoint_x', 'OpPoint_u')]


def stack_bounds(sys):
    LB = np.concatenate([np.asarray(sys['LBx'], float).reshape(len(sys['LBx']), -1),
                         np.asarray(sys['LBu'], float).reshape(len(sys['LBu']), -1)], axis=0)
    UB = np.concatenate([np.asarray(sys['UBx'], float).reshape(len(sys['UBx']), -1),
                         np.asarray(sys['UBu'], float).reshape(len(sys['UBu']), -1)], axis=0)
    if LB.shape[1] == 1:
        LB, UB = LB[:, 0], UB[:, 0]
    return LB, UB
