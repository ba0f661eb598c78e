"""TEST INFRASTRUCTURE.  Dense NumPy twin of the laxMPC FISTA solver, a restatement of the reference's MATLAB twin
platforms/Matlab/spcies_laxMPC_FISTA_solver.m:161-340 (the solver tests/spcies_tester.m:260 compares the generated C against),
including its optional ``lambda`` argument: the dual starting point of a warm start (:161-164, :266-283 -- the warm-up step
evaluates z at ``lambda`` and sets y_0 = lambda_0 = lambda + W^-1 r).  Dense matrices, no structure exploited."""
import numpy as np


def build(sys, param):
    A, B = np.asarray(sys['A'], float), np.asarray(sys['B'], float)
    n, m = B.shape
    N = int(param['N'])
    Q, R, T = (np.asarray(param[k], float) for k in ('Q', 'R', 'T'))
    import scipy.linalg as sla
    H = sla.block_diag(R, np.kron(np.eye(N - 1), sla.block_diag(Q, R)), T)
    Aeq = np.zeros((N * n, N * (n + m)))                       # (u_0, x_1, u_1, ..., x_{N-1}, u_{N-1}, x_N)
    Aeq[:n, :m] = B
    Aeq[:n, m:m + n] = -np.eye(n)
    for l in range(1, N):
        c = m + (l - 1) * (n + m)
        Aeq[l * n:(l + 1) * n, c:c + n + m] = np.hstack([A, B])
        Aeq[l * n:(l + 1) * n, c + n + m:c + 2 * n + m] = -np.eye(n)
    Hinv = np.diag(1.0 / np.diag(H))
    W = Aeq @ Hinv @ Aeq.T
    LB = np.concatenate([sys['LBu'], np.tile(np.concatenate([sys['LBx'], sys['LBu']]), N - 1), sys['LBx']])
    UB = np.concatenate([sys['UBu'], np.tile(np.concatenate([sys['UBx'], sys['UBu']]), N - 1), sys['UBx']])
    return dict(A=A, B=B, n=n, m=m, N=N, Q=Q, R=R, T=T, Aeq=Aeq, Hd=np.diag(H), Wi=np.linalg.inv(W), LB=LB, UB=UB)


def solve(P, x0, xr, ur, lam=None, tol=1e-4, k_max=1000):
    """Returns u_opt, k, e_flag, y (= sol.lambda of the C solver: the dual point of the exit test)."""
    n, m, N = P['n'], P['m'], P['N']
    lam = np.zeros(N * n) if lam is None else np.asarray(lam, float).copy()
    b = np.zeros(N * n)
    b[:n] = -P['A'] @ x0
    q = -np.concatenate([P['R'] @ ur, np.tile(np.concatenate([P['Q'] @ xr, P['R'] @ ur]), N - 1), P['T'] @ xr])
    zof = lambda y: np.clip(-(q - P['Aeq'].T @ y) / P['Hd'], P['LB'], P['UB'])     # solve_boxQP with a diagonal Hessian
    z = zof(lam)
    r = -P['Aeq'] @ z + b
    y = lam + P['Wi'] @ r
    lam_k = y.copy()
    t_k, k = 1.0, 0
    while True:
        k += 1
        t_km1, lam_km1 = t_k, lam_k
        z = zof(y)
        r = -P['Aeq'] @ z + b
        if np.max(np.abs(r)) <= tol:
            return z[:m].copy(), k, 1, y
        if k >= k_max:
            return z[:m].copy(), k, -1, y
        lam_k = P['Wi'] @ r + y
        t_k = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t_km1 ** 2))
        y = lam_k + (t_km1 - 1.0) / t_k * (lam_k - lam_km1)


def closed_loop(P, x0, xr, ur, steps, warm, tol=1e-4, k_max=1000):
    x = np.asarray(x0, float).copy()
    lam = None
    xs, us, ks, es = [x.copy()], [], [], []
    for _ in range(steps):
        u, k, e, y = solve(P, x, xr, ur, lam if warm else None, tol, k_max)
        x = P['A'] @ x + P['B'] @ u
        lam = y if warm != 2 else np.concatenate([y[P['n']:], y[-P['n']:]])      # 2: shifted by one stage
        xs.append(x.copy()); us.append(u); ks.append(k); es.append(e)
    return np.array(xs), np.array(us), np.array(ks), np.array(es)
