"""Named problem configurations.

* ``reference_test(name)``: the settings of the reference's own ``tests/test_<name>.m``
  (3-mass system of tests/spcies_tester.m:90-116, ``tol = 1e-7``, ``k_max = 5000``).
* ``bench_config(name)``: configurations C1..C5 of BASELINE.json / SURVEY.md section 8(d).

Each returns ``dict(sys=..., param=..., kw=...)`` where ``kw`` are the name-value arguments of
``spcies_gen_controller`` (formulation, method, submethod, options).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from . import sysmodel


def _QR(sys):
    p = sys['p']
    return sla.block_diag(15.0 * np.eye(p), np.eye(sys['n'] - p)), 0.1 * np.eye(sys['m'])


def _T_sumP(sys, Q, R):
    """``[~, T] = dlqr(A,B,Q,R); T = diag(sum(T, 2))`` (tests/test_laxMPC_FISTA.m:13-14)."""
    _, P = sysmodel.dlqr(sys['A'], sys['B'], Q, R)
    return np.diag(P.sum(axis=1))


def reference_test(name: str, N: int = 10, masses: int = 3, forces=None, drop: int = 0, **solver_overrides):
    sys = sysmodel.oscillating_masses_sys(p=masses, F=forces, drop=drop)
    Q, R = _QR(sys)
    st = sysmodel.tester_status(sys)
    if name in ('laxMPC_FISTA', 'laxMPC_ADMM'):
        param = dict(Q=Q, R=R, T=_T_sumP(sys, Q, R), N=N)
        so = dict(k_max=5000, tol=1e-7) if name.endswith('FISTA') else dict(rho=15.0, k_max=5000, tol=1e-7)
        kw = dict(formulation='laxMPC', method=name.split('_')[1])
    elif name in ('equMPC_FISTA', 'equMPC_ADMM'):
        param = dict(Q=Q, R=R, N=N)
        so = dict(k_max=5000, tol=1e-7) if name.endswith('FISTA') else dict(rho=15.0, k_max=5000, tol=1e-7)
        kw = dict(formulation='equMPC', method=name.split('_')[1])
    elif name == 'ellipMPC_ADMM':
        param = dict(Q=Q, R=R, T=_T_sumP(sys, Q, R), N=N, P=np.eye(sys['n']), c=st['xr'], r=0.0)
        so = dict(rho=15.0, k_max=5000, tol=1e-7)
        kw = dict(formulation='ellipMPC', method='ADMM')
    elif name == 'ellipMPC_ADMM_soc':
        param = dict(Q=Q, R=R, T=_T_sumP(sys, Q, R), N=N, P=np.eye(sys['n']), c=st['xr'], r=0.0)
        so = dict(rho=15.0, sigma=10.0, k_max=5000, tol_p=1e-7, tol_d=1e-7)
        kw = dict(formulation='ellipMPC', method='ADMM', submethod='soc')
    elif name == 'MPCT_EADMM':
        param = dict(Q=Q, R=R, T=10.0 * Q, S=R, N=N)
        so = dict(rho_base=2.0, rho_mult=20.0, k_max=5000, tol=1e-7)
        kw = dict(formulation='MPCT', method='EADMM')
    elif name == 'MPCT_ADMM_cs':
        # tests/test_MPCT_ADMM.m:6-17 (its rho_base / rho_mult are ignored by this solver: rho stays at the default 1e-2)
        param = dict(Q=Q, R=R, T=10.0 * Q, S=R, N=N)
        so = dict(rho_base=2.0, rho_mult=20.0, k_max=5000, tol=1e-7)
        kw = dict(formulation='MPCT', method='ADMM', submethod='cs')
    elif name == 'MPCT_ADMM_semiband':
        # no test of its own in the reference: the problem of tests/test_MPCT_ADMM.m with a penalty that converges (rho = 2)
        param = dict(Q=Q, R=R, T=10.0 * Q, S=R, N=N)
        so = dict(rho=2.0, k_max=5000, tol_p=1e-7, tol_d=1e-7)
        kw = dict(formulation='MPCT', method='ADMM', submethod='semiband')
    elif name in ('HMPC_ADMM', 'ellipHMPC_ADMM'):
        # tests/test_HMPC_ADMM.m:6-22: method 'ADMM' with the default (empty) submethod = the non-split solver
        param = dict(Q=Q, R=R, N=N, w=3 * 1.627 * 0.2, Te=10.0 * N * Q, Se=R)
        param['Th'] = param['Te']
        param['Sh'] = 0.5 * param['Se']
        so = dict(rho=2.0, sigma=20.0, k_max=5000, tol_p=1e-7, tol_d=1e-7, use_soc=False)
        kw = dict(formulation='HMPC' if name == 'HMPC_ADMM' else 'ellipHMPC', method='ADMM', submethod='')
        if name == 'ellipHMPC_ADMM':
            # coupled constraints y = E x + F u (the formulation needs them); here the box constraints written that way
            n_, m_ = sys['n'], sys['m']
            sys = dict(sys, E=np.vstack([np.eye(n_), np.zeros((m_, n_))]), F=np.vstack([np.zeros((n_, m_)), np.eye(m_)]),
                       LBy=np.concatenate([sys['LBx'], sys['LBu']]), UBy=np.concatenate([sys['UBx'], sys['UBu']]))
            so['sigma'] = 0.0
    elif name in ('HMPC_ADMM_split', 'HMPC_SADMM_split'):
        param = dict(Q=Q, R=R, N=N, w=3 * 1.627 * 0.2, Te=10.0 * N * Q, Se=R)
        param['Th'] = param['Te']
        param['Sh'] = 0.5 * param['Se']
        so = dict(rho=2.0, sigma=20.0, k_max=5000, tol_p=1e-7, tol_d=1e-7, use_soc=False)
        kw = dict(formulation='HMPC', method='SADMM' if 'SADMM' in name else 'ADMM', submethod='split')
    else:
        raise KeyError(name)
    so.update(debug=True, timing=False)
    so.update(solver_overrides)
    kw['options'] = so
    return dict(sys=sys, param=param, kw=kw, status=st, name=name)


def bench_config(name: str):
    """C1/C2: laxMPC FISTA N=10 defaults; C3: equMPC ADMM N=20 rho=15; C4: ellipMPC ADMM_soc N=10;
    C5a: HMPC SADMM_split N=50; C5b: MPCT EADMM N=50."""
    table = {
        'C1': ('laxMPC_FISTA', 10, dict(tol=1e-4, k_max=1000)),
        'C2': ('laxMPC_FISTA', 10, dict(tol=1e-4, k_max=1000)),
        'C3': ('equMPC_ADMM', 20, dict(rho=15.0, tol=1e-4, k_max=1000)),
        'C3f': ('equMPC_ADMM', 20, dict(rho=15.0, tol=1e-4, k_max=1000, precision='float')),
        'C3ff': ('equMPC_ADMM', 20, dict(rho=15.0, tol=1e-4, k_max=1000, precision='float', float_arithmetic=True)),
        'C4': ('ellipMPC_ADMM_soc', 10, dict(rho=15.0, sigma=10.0, tol_p=1e-4, tol_d=1e-4, k_max=1000)),
        'C5a': ('HMPC_SADMM_split', 50, dict(rho=2.0, sigma=20.0, alpha=0.95, tol_p=1e-4, tol_d=1e-4, k_max=1000)),
        'C5b': ('MPCT_EADMM', 50, dict(rho_base=2.0, rho_mult=20.0, tol=1e-4, k_max=1000)),
        # round-2 solvers at the reference tests' problem, default tolerances (rho: the penalty that converges in O(100) iterations)
        'C6': ('MPCT_ADMM_cs', 10, dict(rho=2.0, tol=1e-4, k_max=1000)),
        'C7': ('HMPC_ADMM', 10, dict(rho=2.0, tol_p=1e-4, tol_d=1e-4, k_max=1000)),
        'C8': ('MPCT_ADMM_semiband', 10, dict(rho=2.0, tol_p=1e-4, tol_d=1e-4, k_max=1000)),
    }
    solver, N, so = table[name]
    cfg = reference_test(solver, N=N, **so)
    cfg['kw']['options'].update(debug=False)
    cfg['name'] = name
    cfg['solver'] = solver
    return cfg


def bounds_variant(sys, s: int):
    """Deterministic variant ``s`` of the box constraints (position upper bounds ~ U[0.25, 0.35],
    input lower bounds ~ -U[0.5, 0.8]); used to exercise per-instance bounds."""
    rng = np.random.default_rng(1000 + s)
    sys2 = dict(sys)
    sys2['UBx'] = np.array(sys['UBx'], dtype=float)
    sys2['UBx'][:sys['p']] = rng.uniform(0.25, 0.35, size=sys['p'])
    sys2['LBu'] = -rng.uniform(0.5, 0.8, size=sys['m'])
    return sys2


def per_stage_bounds(sys, N: int):
    """Bounds that tighten along the horizon, one column per stage 0..N (the reference's per-stage bound matrices): the position
    upper bounds shrink from their value to 85 % of it, the input bounds from theirs to 75 %."""
    sys2 = dict(sys)
    w = np.linspace(1.0, 0.0, N + 1)
    p = sys['p']
    UBx = np.repeat(np.asarray(sys['UBx'], float)[:, None], N + 1, axis=1)
    LBx = np.repeat(np.asarray(sys['LBx'], float)[:, None], N + 1, axis=1)
    UBx[:p] *= 0.85 + 0.15 * w
    UBu = np.repeat(np.asarray(sys['UBu'], float)[:, None], N + 1, axis=1) * (0.75 + 0.25 * w)
    LBu = np.repeat(np.asarray(sys['LBu'], float)[:, None], N + 1, axis=1) * (0.75 + 0.25 * w)
    sys2.update(LBx=LBx, UBx=UBx, LBu=LBu, UBu=UBu)
    return sys2
