"""System models used by every benchmark configuration and test.

Host-side restatement of the reference's example-system helpers:

* ``gen_oscillating_masses``  -- +sp_utils/gen_oscillating_masses.m:28-59
* ``example_OscMass``         -- +sp_utils/example_OscMass.m:14-57
* ``tester_system``           -- tests/spcies_tester.m:90-116

MATLAB's ``c2d`` (zero-order hold) and ``dlqr`` are Control System Toolbox calls
whose source is not part of the reference tree; they are restated here with the
textbook formulas (matrix exponential of the augmented matrix; discrete
algebraic Riccati equation).  The discretised ``[A B]`` is pinned against the
15-digit fixture in examples/cl_in_C/main_cl_in_C.c:96 by tests/test_sysmodel.py.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla


def gen_oscillating_masses(M, K, F):
    """Continuous-time model of a chain of masses connected by springs.

    Mirrors +sp_utils/gen_oscillating_masses.m:28-59.  Returns ``(A, B)`` of
    ``xdot = A x + B u`` with state ``[positions; velocities]``.
    """
    M = np.asarray(M, dtype=float).ravel()
    K = np.asarray(K, dtype=float).ravel()
    F = np.asarray(F).ravel().astype(bool)
    p = M.size
    Av = np.zeros((p, p))
    Av[0, 0:2] = [-(K[0] + K[1]), K[1]]
    Av[p - 1, p - 2:p] = [K[p - 1], -(K[p - 1] + K[p])]
    for i in range(1, p - 1):
        Av[i, i - 1:i + 2] = [K[i], -(K[i] + K[i + 1]), K[i + 1]]
    Av = Av / M[:, None]
    A = np.block([[np.zeros((p, p)), np.eye(p)], [Av, np.zeros((p, p))]])
    B = np.vstack([np.zeros((p, p)), np.diag(1.0 / M)])
    B = B[:, F]
    return A, B


def c2d_zoh(A, B, Ts):
    """Zero-order-hold discretisation: expm([[A, B], [0, 0]] * Ts)."""
    n, m = B.shape
    Maug = np.zeros((n + m, n + m))
    Maug[:n, :n] = A
    Maug[:n, n:] = B
    E = sla.expm(Maug * Ts)
    return E[:n, :n].copy(), E[:n, n:].copy()


def dlqr(A, B, Q, R):
    """Discrete LQR: returns ``(K, P)`` like MATLAB's ``[K, S] = dlqr(A,B,Q,R)``."""
    P = sla.solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
    return K, P


def oscillating_masses_sys(Ts: float = 0.2, p: int = 3, F=None, drop: int = 0):
    """The 3-mass system of tests/spcies_tester.m:90-111 (== example_OscMass.m:17-36); ``p`` / ``F`` select another chain of
    +sp_utils/gen_oscillating_masses.m:28-59 (``p`` masses, forces on the masses flagged in ``F``; default: first and last),
    with the same masses / springs / bounds pattern -- the systems of other dimensions used by the shape tests.

    Returns a ``sys`` dict with the reference's field names.
    """
    M = [1.0, 0.5, 1.0] if p == 3 else [1.0 if i % 2 == 0 else 0.5 for i in range(p)]
    K = 2.0 * np.ones(p + 1)
    if F is None:
        F = [1, 0, 1] if p == 3 else [1] + [0] * (p - 2) + [1]
    Ac, Bc = gen_oscillating_masses(M, K, F)
    A, B = c2d_zoh(Ac, Bc, Ts)
    if drop:
        # an odd state dimension for the shape tests: the discrete model without its last `drop` velocity states (not a physical
        # reduction -- just a linear system of that size with the same structure of bounds and weights)
        A, B = A[:-drop, :-drop].copy(), B[:-drop].copy()
    n, m = B.shape
    sys = dict(
        A=A, B=B,
        LBx=-np.concatenate([np.ones(p), 1000.0 * np.ones(p - drop)]),
        UBx=np.concatenate([0.3 * np.ones(p), 1000.0 * np.ones(p - drop)]),
        LBu=-0.8 * np.ones(m),
        UBu=0.8 * np.ones(m),
        p=p, n=n, m=m,
    )
    return sys


def example_OscMass():
    """``[sys, param] = sp_utils.example_OscMass()`` (example_OscMass.m:14-57)."""
    sys = oscillating_masses_sys(0.2)
    p, n, m = sys["p"], sys["n"], sys["m"]
    sys.update(x0=np.zeros(n), u0=np.zeros(m), Nx=np.ones(n), Nu=np.ones(m))
    Q = sla.block_diag(15.0 * np.eye(p), np.eye(p))
    R = 0.1 * np.eye(m)
    _, T = dlqr(sys["A"], sys["B"], Q, R)
    param = dict(Q=Q, R=R, T=T, N=10)
    return sys, param


def tester_status(sys):
    """The fixed test point of tests/spcies_tester.m:114-116."""
    n, m = sys["n"], sys["m"]
    x = 0.02 * np.ones(n)
    ur = 0.5 * np.ones(m)
    xr = steady_state(sys, ur)
    return dict(x=x, ur=ur, xr=xr)


def steady_state(sys, ur):
    """``xr = (A - I) \\ (-B ur)`` (tests/spcies_tester.m:116); ``ur`` may be ``[B, m]``."""
    A, B = sys["A"], sys["B"]
    n = A.shape[0]
    ur = np.asarray(ur, dtype=float)
    rhs = -(B @ ur.T) if ur.ndim == 2 else -(B @ ur)
    xr = np.linalg.solve(A - np.eye(n), rhs)
    return xr.T.copy() if ur.ndim == 2 else xr


def synthetic_batch(sys, B, seed=0, with_r=False, chunk=1 << 20):
    """Seeded synthetic batch of SURVEY.md section 8(d).

    ``x0 ~ U[-0.25, 0.25]^n``, ``ur ~ U[-0.6, 0.6]^m``, ``xr`` the matching steady
    state; optionally ``r ~ U[0.05, 0.5]`` (ellipsoid size, config C4).
    Returned arrays are C-contiguous, instance-major (``[B, n]``).
    """
    n, m = sys["n"], sys["m"]
    rng = np.random.default_rng(seed)
    x0 = np.empty((B, n))
    ur = np.empty((B, m))
    r = np.empty(B) if with_r else None
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        x0[s:e] = rng.uniform(-0.25, 0.25, size=(e - s, n))
        ur[s:e] = rng.uniform(-0.6, 0.6, size=(e - s, m))
        if with_r:
            r[s:e] = rng.uniform(0.05, 0.5, size=e - s)
    xr = steady_state(sys, ur)
    out = dict(x0=x0, xr=np.ascontiguousarray(xr), ur=ur)
    if with_r:
        out["r"] = r
    return out


def perturbed_models(sys, param, B, seed=0, rel=0.03):
    """Per-instance models for the TIME_VARYING solvers: ``A_i, B_i`` = the nominal model with every entry perturbed by up to
    ``rel`` (relative), ``Q_i, R_i`` = the nominal diagonal weights scaled by U[0.5, 2], bounds with the position upper bounds in
    U[0.25, 0.35].  Returns ``(A [B,n,n], Bm [B,n,m], Q [B,n], R [B,m]), LB [B,nm], UB [B,nm]``."""
    rng = np.random.default_rng(seed)
    n, m, p = sys['n'], sys['m'], sys['p']
    A = sys['A'][None] * (1.0 + rel * rng.uniform(-1, 1, (B, n, n)))
    Bm = sys['B'][None] * (1.0 + rel * rng.uniform(-1, 1, (B, n, m)))
    Q = np.diag(param['Q'])[None] * rng.uniform(0.5, 2.0, (B, n))
    R = np.diag(param['R'])[None] * rng.uniform(0.5, 2.0, (B, m))
    LB = np.tile(np.concatenate([sys['LBx'], sys['LBu']]), (B, 1))
    UB = np.tile(np.concatenate([sys['UBx'], sys['UBu']]), (B, 1))
    UB[:, :p] = rng.uniform(0.25, 0.35, (B, p))
    return (np.ascontiguousarray(A), np.ascontiguousarray(Bm), Q, R), LB, UB
