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


def check_glpk_basis(s, slots_t, days, k, x, row, col):
    """`row` [m] / `col` [n]: GLPK-coded statuses of scenario k (BasisManager.row_stat / col_stat, cython_glpk.pyx:197-232) for
    the LP of this timestep, `x` the engine's column values.  Exactly m variables are basic, every non-basic variable sits on
    the bound its status names, and the basic columns / rows form a non-singular basis matrix whose solution is x."""
    BS, NL, NU, NF, NS = 1, 2, 3, 4, 5
    m, n = s.n_rows, s.n_cols
    A, lo, hi, c = assemble(s, slots_t, k, days)
    A = np.asarray(A.todense()) if hasattr(A, "todense") else np.asarray(A)
    r = A @ x
    assert (row == BS).sum() + (col == BS).sum() == m, ((row == BS).sum(), (col == BS).sum(), m)
    scale = max(1.0, np.abs(x).max())
    for j in range(n):
        if col[j] in (NL, NS):
            assert abs(x[j]) <= 1e-7 * scale, (j, x[j])
        assert col[j] != NU            # original columns are x >= 0 without an upper bound
    for i in range(m):
        if row[i] == NL or row[i] == NS:
            assert abs(r[i] - lo[i]) <= 1e-7 * scale, (i, r[i], lo[i])
        elif row[i] == NU:
            assert abs(r[i] - hi[i]) <= 1e-7 * scale, (i, r[i], hi[i])
    # basis matrix of [A | -I] v = 0: basic columns of A and -e_i of the basic rows; non-singular, and solving with the
    # non-basic variables on their bounds reproduces x
    bc = np.nonzero(col == BS)[0]
    br = np.nonzero(row == BS)[0]
    B = np.hstack([A[:, bc], -np.eye(m)[:, br]])
    assert np.linalg.matrix_rank(B) == m
    rhs = np.zeros(m)
    for i in range(m):
        if row[i] != BS:
            rhs[i] += (hi[i] if row[i] == NU else (0.0 if row[i] == NF else lo[i]))
    sol = np.linalg.solve(B, rhs)      # A_B x_B - r_B = r_N  (non-basic columns are at 0)
    assert np.allclose(sol[: len(bc)], x[bc], rtol=1e-7, atol=1e-7 * scale)
