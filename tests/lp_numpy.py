"""numpy restatement of the per-scenario LP assembly (bounds, costs, dynamic coefficients) used by the
tests to hand the very same LP to scipy's HiGHS as an independent check of the oracle / the CUDA
engine.  Follows pywr/solvers/cython_glpk.pyx:66-95 (typing), :880-897 (costs), :903-926 (factors),
:932-1019 (bounds).  Test infrastructure."""
import numpy as np

REF_STATE = -(2**31)


def ref_value(s, ref, slots, k):
    return slots[ref, k] if ref >= 0 else s.consts[-ref - 1]


def type_bounds(lo, hi):
    if abs(lo) < 1e-8:
        lo = 0.0
    if abs(hi) < 1e-8:
        hi = 0.0
    if abs(lo - hi) < 1e-8:
        hi = lo
    return lo, hi


def assemble(s, slots, k, days, vol_state=None):
    """Returns (A dense, row_lo, row_hi, cost) of scenario k."""
    m, n = s.n_rows, s.n_cols
    A = s.dense()
    for q in range(len(s.dyn_a_nz)):
        nz = s.dyn_a_nz[q]
        row = int(np.searchsorted(s.a_indptr, nz, side="right") - 1)
        A[row, s.a_col[nz]] += s.dyn_a_scale[q] * ref_value(s, s.dyn_a_ref[q], slots, k) - s.a_val[nz]
    lo = np.zeros(m)
    hi = np.zeros(m)
    for i in range(m):
        if s.row_kind[i] == 0:
            l, u = type_bounds(ref_value(s, s.row_lb_ref[i], slots, k), ref_value(s, s.row_ub_ref[i], slots, k))
        else:
            q = s.row_lb_ref[i]
            maxv = ref_value(s, s.st_maxv_ref[q], slots, k)
            minv = ref_value(s, s.st_minv_ref[q], slots, k)
            vol = vol_state[q, k] if s.st_vol_ref[q] == REF_STATE else ref_value(s, s.st_vol_ref[q], slots, k)
            active = True
            if s.st_kind[q] == 1:
                active = ref_value(s, s.st_active_ref[q], slots, k) != 0.0
            if maxv == minv:
                l, u = 0.0, 0.0
            elif not active:
                l, u = -np.inf, np.inf
            else:
                l, u = type_bounds(-max(vol - minv, 0.0) / days, max(maxv - vol, 0.0) / days)
        lo[i], hi[i] = l, u
    c = np.zeros(n)
    for j in range(n):
        acc = 0.0
        for t in range(s.cost_indptr[j], s.cost_indptr[j + 1]):
            acc += s.cost_w[t] * ref_value(s, s.cost_ref[t], slots, k)
        if s.cost_clip and abs(acc) < 1e-8:
            acc = 0.0
        c[j] = acc
    return A, lo, hi, c


def highs_objective(s, slots, k, days, vol_state=None):
    """Optimal objective of scenario k from scipy HiGHS (dual simplex, presolve off), or None."""
    from scipy.optimize import linprog

    A, lo, hi, c = assemble(s, slots, k, days, vol_state)
    eq = lo == hi
    A_ub, b_ub = [], []
    for i in np.nonzero(~eq)[0]:
        if np.isfinite(hi[i]):
            A_ub.append(A[i]); b_ub.append(hi[i])
        if np.isfinite(lo[i]):
            A_ub.append(-A[i]); b_ub.append(-lo[i])
    res = linprog(c, A_ub=np.array(A_ub) if A_ub else None, b_ub=np.array(b_ub) if b_ub else None,
                  A_eq=A[eq] if eq.any() else None, b_eq=lo[eq] if eq.any() else None,
                  bounds=(0, None), method="highs-ds", options={"presolve": False})
    return res.fun if res.status == 0 else None


def residuals(s, slots, k, days, x, vol_state=None):
    """max bound violation of a column vector x for scenario k"""
    A, lo, hi, c = assemble(s, slots, k, days, vol_state)
    r = A @ x
    return max(0.0, float(np.max(lo - r)), float(np.max(r - hi)), float(np.max(-x))), float(c @ x)
