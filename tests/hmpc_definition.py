"""TEST INFRASTRUCTURE.  The harmonic-MPC problem written down from its *definition* (Krupa, Limon, Alamo, "Harmonic based
model predictive control for set-point tracking", IEEE TAC 2022; the formulation documented in the reference's
docs/HMPC.md), independently of the block formulas H11..H33 / G of compute_HMPC_ADMM_split_ingredients.m:71-142 that the
product's spcies_b200/formulations/HMPC.py restates:

    min  sum_{j=0}^{N-1} |x_j - xh_j|_Q^2 + |u_j - uh_j|_R^2 + |x_e - xr|_Te^2 + |x_s|_Th^2 + |x_c|_Th^2
                                                        + |u_e - ur|_Se^2 + |u_s|_Sh^2 + |u_c|_Sh^2            (all times 1/2)
    with the harmonic artificial reference  xh_j = x_e + x_s sin(w j) + x_c cos(w j)  (same for u),
    s.t. x_0 = x(t), x_{j+1} = A x_j + B u_j, x_N = xh_N,
         x_e = A x_e + B u_e,  x_s cos w - x_c sin w = A x_s + B u_s,  x_s sin w + x_c cos w = A x_c + B u_c,
         box bounds on u_0, (x_j, u_j)_{j=1..N-1}, and  LB_i <= y_e,i -+ |(y_s,i, y_c,i)| <= UB_i  for every signal i of y = (x, u).

Decision vector (the reference's order): z = (u_0, x_1, u_1, ..., x_{N-1}, u_{N-1}, x_e, x_s, x_c, u_e, u_s, u_c).

``quadratic_from_definition`` evaluates the cost / the equality residuals as *functions* and recovers H, q, G, b by probing them
with unit vectors; ``admm_twin`` is a dense NumPy ADMM on those matrices (the role of the reference's MATLAB twin solvers,
platforms/Matlab/spcies_HMPC_*_solver.m, tests/spcies_tester.m:260).
"""
import numpy as np


def unpack(z, n, m, N):
    nm = n + m
    u = [z[0:m]]
    x = [None]
    for j in range(1, N):
        o = m + (j - 1) * nm
        x.append(z[o:o + n])
        u.append(z[o + n:o + nm])
    o = m + (N - 1) * nm
    xe, xs, xc = z[o:o + n], z[o + n:o + 2 * n], z[o + 2 * n:o + 3 * n]
    o += 3 * n
    ue, us, uc = z[o:o + m], z[o + m:o + 2 * m], z[o + 2 * m:o + 3 * m]
    return x, u, xe, xs, xc, ue, us, uc


def cost(z, x0, xr, ur, sys, param):
    n, m, N, w = sys['n'], sys['m'], int(param['N']), float(param['w'])
    Q, R, Te, Th, Se, Sh = (np.asarray(param[k], float) for k in ('Q', 'R', 'Te', 'Th', 'Se', 'Sh'))
    x, u, xe, xs, xc, ue, us, uc = unpack(z, n, m, N)
    x[0] = np.asarray(x0, float)
    f = 0.0
    for j in range(N):
        dx = x[j] - xe - xs * np.sin(w * j) - xc * np.cos(w * j)
        du = u[j] - ue - us * np.sin(w * j) - uc * np.cos(w * j)
        f += 0.5 * dx @ Q @ dx + 0.5 * du @ R @ du
    f += 0.5 * (xe - xr) @ Te @ (xe - xr) + 0.5 * xs @ Th @ xs + 0.5 * xc @ Th @ xc
    f += 0.5 * (ue - ur) @ Se @ (ue - ur) + 0.5 * us @ Sh @ us + 0.5 * uc @ Sh @ uc
    return f


def equality_residual(z, x0, sys, param):
    """Rows in the reference's order: dynamics j = 0..N-2, terminal x_N = xh_N written through the last dynamics row, then the
    three harmonic steady-state conditions; sign: [A B] (x_j, u_j) - x_{j+1}."""
    n, m, N, w = sys['n'], sys['m'], int(param['N']), float(param['w'])
    A, B = np.asarray(sys['A'], float), np.asarray(sys['B'], float)
    x, u, xe, xs, xc, ue, us, uc = unpack(z, n, m, N)
    x[0] = np.asarray(x0, float)
    r = []
    for j in range(N - 1):
        r.append(A @ x[j] + B @ u[j] - x[j + 1])
    xN = xe + xs * np.sin(w * N) + xc * np.cos(w * N)
    r.append(A @ x[N - 1] + B @ u[N - 1] - xN)
    r.append(A @ xe + B @ ue - xe)
    r.append(A @ xs + B @ us - (xs * np.cos(w) - xc * np.sin(w)))
    r.append(A @ xc + B @ uc - (xs * np.sin(w) + xc * np.cos(w)))
    return np.concatenate(r)


def quadratic_from_definition(x0, xr, ur, sys, param):
    n, m, N = sys['n'], sys['m'], int(param['N'])
    dim = (N - 1) * (n + m) + m + 3 * (n + m)
    I = np.eye(dim)
    f0 = cost(np.zeros(dim), x0, xr, ur, sys, param)
    f1 = np.array([cost(I[i], x0, xr, ur, sys, param) for i in range(dim)])
    H = np.zeros((dim, dim))
    for i in range(dim):
        for j in range(i, dim):
            H[i, j] = H[j, i] = cost(I[i] + I[j], x0, xr, ur, sys, param) - f1[i] - f1[j] + f0
    q = f1 - f0 - 0.5 * np.diag(H)
    g0 = equality_residual(np.zeros(dim), x0, sys, param)
    G = np.stack([equality_residual(I[i], x0, sys, param) - g0 for i in range(dim)], axis=1)
    return H, q, G, -g0          # G z = b


def proj_diamond(y, lb, ub):
    """Projection of (y_e, y_s, y_c) onto  lb <= y_e - |(y_s, y_c)|,  y_e + |(y_s, y_c)| <= ub  (two shifted cones in sequence,
    +sp_utils/proj_D.m)."""
    def ssoc(x, alpha, d):
        x0, nx = x[0], np.hypot(x[1], x[2])
        if nx <= alpha * (x0 - d):
            return x
        if nx <= -alpha * (x0 - d):
            return np.array([d, 0.0, 0.0])
        s = 0.5 * (alpha * (x0 - d) + nx)
        return np.array([s * alpha + d, s * x[1] / nx, s * x[2] / nx])
    return ssoc(ssoc(np.asarray(y, float), 1.0, lb), -1.0, ub)


def admm_twin(H, q, G, b, sys, param, rho=2.0, tol=1e-10, k_max=200000):
    """min 1/2 z'Hz + q'z  s.t.  G z = b,  z in Z  by plain ADMM (z = v splitting), dense NumPy."""
    n, m, N = sys['n'], sys['m'], int(param['N'])
    nm = n + m
    dim = H.shape[0]
    nbox = dim - 3 * nm
    LB = np.concatenate([sys['LBu'], np.tile(np.concatenate([sys['LBx'], sys['LBu']]), N - 1)])
    UB = np.concatenate([sys['UBu'], np.tile(np.concatenate([sys['UBx'], sys['UBu']]), N - 1)])
    LBy, UBy = np.concatenate([sys['LBx'], sys['LBu']]), np.concatenate([sys['UBx'], sys['UBu']])
    ne = G.shape[0]
    K = np.block([[H + rho * np.eye(dim), G.T], [G, np.zeros((ne, ne))]])
    Ki = np.linalg.inv(K)
    v = np.zeros(dim)
    lam = np.zeros(dim)
    o = nbox
    idx = lambda i: ([o + i, o + n + i, o + 2 * n + i] if i < n else
                     [o + 3 * n + (i - n), o + 3 * n + m + (i - n), o + 3 * n + 2 * m + (i - n)])
    for k in range(1, k_max + 1):
        rhs = np.concatenate([rho * v - lam - q, b])
        z = (Ki @ rhs)[:dim]
        vo = v
        t = z + lam / rho
        v = t.copy()
        v[:nbox] = np.clip(t[:nbox], LB, UB)
        for i in range(nm):
            ii = idx(i)
            v[ii] = proj_diamond(t[ii], LBy[i], UBy[i])
        lam = lam + rho * (z - v)
        if max(np.max(np.abs(z - v)), np.max(np.abs(v - vo))) <= tol:
            return v, k, True
    return v, k_max, False
