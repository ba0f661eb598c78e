"""Host-side mirrors of the reference's ``+sp_utils`` sparse helpers and projections.

These run offline (when a solver is generated) exactly like their MATLAB twins;
the on-line versions are device functions in ``csrc/spcies_sparse.cuh`` and
``csrc/spcies_proj.cuh``.

Index arrays are **0-based** here (the reference builds them 1-based and shifts
by one on emission, e.g. cons_ellipMPC_ADMM_soc_C.m:99-100).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla


@dataclass
class Sparse:
    """CSR (``row`` = row pointers, ``col`` = column indices) or CSC
    (``col`` = column pointers, ``row`` = row indices) container, like the
    struct returned by full2CSR.m / full2CSC.m."""
    val: np.ndarray
    col: np.ndarray
    row: np.ndarray
    nnz: int
    nrow: int
    ncol: int


def full2CSR(M, threshold=0.0) -> Sparse:
    """Dense -> CSR by a row-major scan keeping ``abs(M) > threshold``
    (+sp_utils/full2CSR.m:28-62)."""
    M = np.asarray(M, dtype=float)
    n, m = M.shape
    mask = np.abs(M) > threshold
    val = M[mask]                       # boolean indexing scans row-major
    col = np.nonzero(mask)[1].astype(np.int32)
    row = np.zeros(n + 1, dtype=np.int32)
    row[1:] = np.cumsum(mask.sum(axis=1))
    return Sparse(val=val.copy(), col=col, row=row, nnz=int(val.size), nrow=n, ncol=m)


def full2CSC(M, threshold=0.0) -> Sparse:
    """Dense -> CSC as the CSR of the transpose with names swapped
    (+sp_utils/full2CSC.m:25-44)."""
    M = np.asarray(M, dtype=float)
    n, m = M.shape
    t = full2CSR(M.T, threshold)
    return Sparse(val=t.val, row=t.col, col=t.row, nnz=t.nnz, nrow=n, ncol=m)


def chol_upper(M):
    """MATLAB ``chol``: the upper-triangular factor ``R`` with ``R' R = M``."""
    return sla.cholesky(np.asarray(M, dtype=float), lower=False)


def full2LDL(M, for_LDLsolve=False):
    """LDL' from the Cholesky factor (+sp_utils/full2LDL.m:16-57):
    ``L = R' diag(1/diag R)``, ``D = diag(R)^2``.  With ``for_LDLsolve`` returns
    ``(L_val, L_row, L_colptr, Dinv)`` of the CSC form of ``L - I``."""
    Mc = chol_upper(M)
    d = np.diag(Mc)
    L = Mc.T @ np.diag(1.0 / d)
    D = np.diag(d ** 2)
    if not for_LDLsolve:
        return L, D
    L_CSC = full2CSC(L - np.eye(L.shape[0]))
    Dinv = 1.0 / np.diag(D)
    return L_CSC.val, L_CSC.row, L_CSC.col, Dinv


def smv(val, col, row_ptr, x):
    """CSR sparse matrix-vector product (+sp_utils/smv.m:23-36)."""
    n = len(row_ptr) - 1
    y = np.zeros(n)
    for i in range(n):
        acc = 0.0
        for j in range(row_ptr[i], row_ptr[i + 1]):
            acc = acc + val[j] * x[col[j]]
        y[i] = acc
    return y


def LDLsolve(val, row, col_ptr, Dinv, b):
    """QDLDL-style solve of ``L D L' x = b`` with CSC ``L - I``
    (+sp_utils/LDLsolve.m:22-49)."""
    x = np.array(b, dtype=float)
    n = x.size
    for i in range(n):
        value = x[i]
        for j in range(col_ptr[i], col_ptr[i + 1]):
            x[row[j]] -= val[j] * value
    x *= Dinv
    for i in range(n - 1, -1, -1):
        value = x[i]
        for j in range(col_ptr[i], col_ptr[i + 1]):
            value -= val[j] * x[row[j]]
        x[i] = value
    return x


def proj_SOC(x):
    """Projection onto ``||x[1:]|| <= x[0]`` (+sp_utils/proj_SOC.m:12-27)."""
    x = np.asarray(x, dtype=float)
    x0 = x[0]
    nx1 = np.linalg.norm(x[1:], 2)
    if nx1 <= x0:
        return x.copy()
    if nx1 <= -x0:
        return np.zeros_like(x)
    return 0.5 * (x0 + nx1) * np.concatenate([[1.0], x[1:] / nx1])


def proj_SSOC(x, alpha, d):
    """Projection onto the shifted cone ``||x[1:]|| <= alpha (x[0] - d)``
    (+sp_utils/proj_SSOC.m:14-29)."""
    x = np.asarray(x, dtype=float)
    x0 = x[0]
    nx1 = np.linalg.norm(x[1:], 2)
    shift = np.concatenate([[d], np.zeros(x.size - 1)])
    if nx1 <= alpha * (x0 - d):
        return x.copy()
    if nx1 <= -alpha * (x0 - d):
        return shift
    return 0.5 * (alpha * (x0 - d) + nx1) * np.concatenate([[alpha], x[1:] / nx1]) + shift


def proj_D(x, lb, ub):
    """Projection onto the 'diamond' set = SSOC(+1, lb) then SSOC(-1, ub)
    (+sp_utils/proj_D.m:19-23)."""
    return proj_SSOC(proj_SSOC(x, 1.0, lb), -1.0, ub)
