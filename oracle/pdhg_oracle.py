"""CPU ORACLE of the batched restarted-PDHG path (large networks).  TEST INFRASTRUCTURE ONLY.

numpy / scipy.sparse restatement, one scenario at a time, of what ``pywr_b200/csrc/pdhg_engine.cuh``
does for all scenarios at once.  Only tests/, bench.py's cpu legs and ``__graft_entry__.smoke()`` may
import this module; the product package never does.

What it replaces in the reference: the same per-scenario LP of ``CythonGLPKEdgeSolver._solve_scenario``
(pywr/solvers/cython_glpk.pyx:1649-1918; bounds typing :66-95,:932-1019, costs :880-897) for networks
whose dictionary does not fit the simplex kernels.  GLPK's simplex is replaced by a first-order method
(BASELINE.json north_star), so parity is by objective (1e-7 relative) and feasibility residuals, pinned
against scipy HiGHS on the committed golden LPs (tests/golden/*.pdhg.npz).

Algorithm (identical on both sides; every step below has the same name in the CUDA file):
  1. structural presolve, once per structure: rows with a single non-zero become column bounds
     (most non-storage rows of the edge form, cython_glpk.pyx:1300-1312), rows that can never bind
     (non-negative coefficients on x >= 0, constant bounds lb <= 0, ub = +inf) are dropped;
  2. Ruiz (10 sweeps) + Pock-Chambolle (alpha = 1) diagonal scaling of the reduced matrix and a
     power-iteration estimate of its 2-norm: step sizes tau = eta / omega, sigma = eta * omega,
     eta = 0.99 / ||A||;
  3. reflected restarted Halpern PDHG (r2HPDHG):  z+ = w (2 T(z) - z) + (1 - w) z_anchor,
     w = (k+1)/(k+2), T = one PDHG step
         x' = clip(x - tau (c - A^T y), lc, uc)
         y' = v - clip(v, -sigma hi, -sigma lo),   v = y - sigma A (2 x' - x)
     every ``check`` iterations: KKT residuals of T(z); converged -> stop; restart (anchor := T(z),
     primal weight omega updated from the movement since the last restart if both iterates moved by more
     than 1e-5 of their norm; if only y moved and the primal residual dominates (x pinned at its bounds)
     omega *= 4, if only x moved and the primal residual does not dominate omega /= 4; always within 1e4 of
     ||c||/||b||) when the fixed-point
     residual fell to 20 %, or to 80 % and is growing again, or after 36 % of all iterations;
  4. warm start: every scenario keeps (x, y, omega) from its previous timestep (role of the carried-over
     basis, cython_glpk.pyx:197-232);
  5. bound snapping of the columns and un-scaling.
"""

import numpy as np
import scipy.sparse as sp

REF_STATE = -(2 ** 31)
BIG = 1e300
ST_OPTIMAL, ST_INFEASIBLE, ST_ITERLIMIT, ST_NAN = 0, 1, 3, 4
OMEGA_MOVE, OMEGA_SPAN, OMEGA_KICK = 1e-5, 1e4, 4.0


def _fin(a):
    return np.where(np.abs(a) < BIG, a, 0.0)


class PdhgPlan:
    """Steps 1 and 2 (host side, shared by all scenarios)."""

    RUIZ_SWEEPS = 10
    AGG_LEN = 12
    POWER_ITERS = 300

    def __init__(self, s):
        if len(s.dyn_a_nz):
            raise ValueError("the PDHG path does not support per-scenario matrix coefficients")
        self.s = s
        m, n = s.n_rows, s.n_cols
        A = sp.csr_matrix((s.a_val, s.a_col, s.a_indptr), shape=(m, n))
        A.sum_duplicates()
        self.A_full = A
        consts = np.asarray(s.consts, float)

        def const_of(ref):
            return consts[-ref - 1] if ref < 0 else None

        # rows as sorted (column, value) lists without zeros
        rows = []
        for i in range(m):
            k0, k1 = A.indptr[i], A.indptr[i + 1]
            rows.append([(int(j), float(v)) for j, v in sorted(zip(A.indices[k0:k1], A.data[k0:k1])) if v != 0.0])
        # equality-row elimination (PdhgPlan::build in pdhg_engine.cuh, same order of decisions and of floating-point operations),
        # two rules applied until neither fires:
        #  (1) doubletons: a row  a x_p + b x_q = 0  with constant zero bounds and a b < 0 (the mass balance of a link with one edge
        #      in and one out) fixes x_q = rho x_p, rho = -a / b > 0: column q is folded into column p (in every other row, in the
        #      cost and - through its singleton rows - in the bounds) and the row disappears.  Chains of links collapse.
        #  (2) confluences: a row  b x_q + sum a_k x_k = 0  (3..AGG_LEN entries, constant zero bounds) where x_q is the only entry of
        #      its sign fixes x_q = sum rho_k x_k, rho_k = -a_k / b > 0; x_q >= 0 is implied, so q goes if every singleton row on
        #      it can never bind.  q is replaced by the combination in every other row.
        elim_row = np.full(n, -1, np.int32)
        gone = np.zeros(m, bool)
        comb = [[] for _ in range(n)]           # folded column -> [(surviving column, ratio)]
        colrows = [set() for _ in range(n)]     # rows (gone or not) that list the column
        for i in range(m):
            for j, _ in rows[i]:
                colrows[j].add(i)
        combcols = [set() for _ in range(n)]    # folded columns whose combination lists the column

        def zero_bounds(i):
            if s.row_kind[i] != 0:
                return False
            lb, ub = const_of(int(s.row_lb_ref[i])), const_of(int(s.row_ub_ref[i]))
            return lb is not None and ub is not None and abs(lb) < 1e-8 and abs(ub) < 1e-8

        def never_binds(i):
            r = rows[i]
            if s.row_kind[i] != 0 or not r or any(v < 0.0 for _, v in r):
                return False
            lb, ub = const_of(int(s.row_lb_ref[i])), const_of(int(s.row_ub_ref[i]))
            return lb is not None and ub is not None and lb < 1e-8 and ub >= BIG

        def substitute(lst, q, C):
            """entry q (coefficient v) of the sorted list replaced by v * C, merged into the entries already there"""
            d = dict(lst)
            if q not in d:
                return lst
            v = d.pop(q)
            for col, rho in C:
                nv = d.get(col, 0.0) + v * rho
                if nv != 0.0:
                    d[col] = nv
                elif col in d:
                    del d[col]
            return sorted(d.items())

        def eliminate(i, q, C):
            for k in sorted(colrows[q]):
                if k == i or gone[k]:
                    continue
                rows[k] = substitute(rows[k], q, C)
                for col, _ in C:
                    colrows[col].add(k)         # a superset is enough: membership is re-checked where it matters
            for j in sorted(combcols[q]):
                comb[j] = substitute(comb[j], q, C)
                for col, _ in C:
                    combcols[col].add(j)
            combcols[q] = set()
            comb[q] = list(C)
            for col, _ in C:
                combcols[col].add(q)
            gone[i] = True
            elim_row[q] = i

        import os
        rule1 = not os.environ.get("B200LP_NO_DOUBLETON")
        rule2 = rule1 and not os.environ.get("B200LP_NO_AGGREGATE")
        changed = rule1
        while changed:
            changed = False
            ccount = np.zeros(n, np.int64)
            for i in range(m):
                if not gone[i]:
                    for j, _ in rows[i]:
                        ccount[j] += 1
            for i in range(m):
                r = rows[i]
                if gone[i] or len(r) != 2 or not zero_bounds(i):
                    continue
                (p0, a), (p1, b) = r
                if a * b >= 0.0:
                    continue
                q = p0 if (ccount[p0] < ccount[p1] or (ccount[p0] == ccount[p1] and p0 > p1)) else p1
                pcol = p1 if q == p0 else p0
                aq, ap = (a, b) if q == p0 else (b, a)
                eliminate(i, q, [(pcol, -ap / aq)])
                ccount[q] = 0
                ccount[pcol] = sum(1 for k in colrows[pcol] if not gone[k] and any(j == pcol for j, _ in rows[k]))
                changed = True
            if not rule2:
                continue
            for i in range(m):
                r = rows[i]
                if gone[i] or len(r) < 3 or len(r) > self.AGG_LEN or not zero_bounds(i):
                    continue
                pos = [t for t, (_, v) in enumerate(r) if v > 0.0]
                if len(pos) == 1:
                    at = pos[0]
                elif len(pos) == len(r) - 1:
                    at = [t for t, (_, v) in enumerate(r) if not v > 0.0][0]
                else:
                    continue
                q, b = r[at]
                if any(k != i and not gone[k] and len(rows[k]) == 1 and rows[k][0][0] == q and not never_binds(k) for k in colrows[q]):
                    continue
                eliminate(i, q, [(col, -v / b) for t, (col, v) in enumerate(r) if t != at])
                changed = True
        alias_root = np.arange(n)
        for j in range(n):
            if comb[j]:
                alias_root[j] = comb[j][0][0]       # != j: marks the column as folded
        self.alias_root, self.elim_row = alias_root, elim_row
        self.comb = [comb[j] if comb[j] else [(j, 1.0)] for j in range(n)]
        self.members = [[] if comb[j] else [(j, 1.0)] for j in range(n)]
        for q in range(n):
            for col, ratio in comb[q]:
                self.members[col].append((q, ratio))
        keep, sing_row, sing_col, sing_coef = [], [], [], []
        for i in range(m):
            if gone[i]:
                continue
            r = rows[i]
            cnt = len(r)
            if cnt == 1:
                sing_row.append(i); sing_col.append(r[0][0]); sing_coef.append(r[0][1])
                continue
            if s.row_kind[i] == 0 and cnt > 0 and all(v >= 0.0 for _, v in r):
                lb, ub = const_of(int(s.row_lb_ref[i])), const_of(int(s.row_ub_ref[i]))
                if lb is not None and ub is not None and lb < 1e-8 and ub >= BIG:
                    continue  # can never bind
            keep.append(i)
        # the transformed matrix (folded columns) replaces A from here on
        ri, ci, vi = [], [], []
        for i in range(m):
            if not gone[i]:
                for j, v in rows[i]:
                    ri.append(i); ci.append(j); vi.append(v)
        A = sp.csr_matrix((vi, (ri, ci)), shape=(m, n))
        self.keep = np.array(keep, np.int32)
        self.sing_row = np.array(sing_row, np.int32)
        self.sing_col = np.array(sing_col, np.int32)
        self.sing_coef = np.array(sing_coef, float)
        self.m_r, self.n = len(keep), n
        Ar = A[self.keep].tocsr() if len(keep) else sp.csr_matrix((0, n))
        dr, dc = np.ones(self.m_r), np.ones(n)
        B = Ar.copy().astype(float)
        for _ in range(self.RUIZ_SWEEPS):
            rn = np.sqrt(abs(B).max(axis=1).toarray().ravel()) if self.m_r else np.ones(0)
            cn = np.sqrt(abs(B).max(axis=0).toarray().ravel()) if self.m_r else np.ones(n)
            rn[rn == 0] = 1.0; cn[cn == 0] = 1.0
            B = sp.diags(1.0 / rn) @ B @ sp.diags(1.0 / cn)
            dr /= rn; dc /= cn
        rn = np.sqrt(np.asarray(abs(B).sum(axis=1)).ravel()) if self.m_r else np.ones(0)
        cn = np.sqrt(np.asarray(abs(B).sum(axis=0)).ravel()) if self.m_r else np.ones(n)
        rn[rn == 0] = 1.0; cn[cn == 0] = 1.0
        dr /= rn; dc /= cn
        self.dr, self.dc = dr, dc
        self.A = (sp.diags(dr) @ Ar @ sp.diags(dc)).tocsr()
        self.AT = self.A.T.tocsr()
        v = np.ones(n) / np.sqrt(n)
        nv = 0.0
        for _ in range(self.POWER_ITERS):
            v = self.AT @ (self.A @ v)
            nv = np.linalg.norm(v)
            if nv == 0.0:
                break
            v /= nv
        self.norm_a = np.sqrt(nv) if nv > 0 else 1.0
        self.eta = 0.99 / self.norm_a


def _ref_value(s, ref, slots_k):
    return slots_k[ref] if ref >= 0 else s.consts[-ref - 1]


def _clip8(v):
    return 0.0 if abs(v) < 1e-8 else v


def row_bounds(s, i, slots_k, days, vol_k):
    """(lb, ub) of original row i (cython_glpk.pyx:1753-1840, same typing as the route form)."""
    if s.row_kind[i] == 0:
        lo, hi = _clip8(_ref_value(s, s.row_lb_ref[i], slots_k)), _clip8(_ref_value(s, s.row_ub_ref[i], slots_k))
    else:
        q = s.row_lb_ref[i]
        maxv, minv = _ref_value(s, s.st_maxv_ref[q], slots_k), _ref_value(s, s.st_minv_ref[q], slots_k)
        vol = vol_k[q] if s.st_vol_ref[q] == REF_STATE else _ref_value(s, s.st_vol_ref[q], slots_k)
        active = True if s.st_kind[q] == 0 else _ref_value(s, s.st_active_ref[q], slots_k) != 0.0
        if maxv == minv:
            return 0.0, 0.0
        if not active:
            return -np.inf, np.inf
        lo, hi = _clip8(-max(vol - minv, 0.0) / days), _clip8(max(maxv - vol, 0.0) / days)
    if abs(lo - hi) < 1e-8:
        hi = lo
    return lo, hi


def assemble(plan, slots_k, days, vol_k):
    """Scaled per-scenario data of the reduced LP: lo, hi [m_r], c, lc, uc [n]; status."""
    s = plan.s
    n = plan.n
    lo = np.empty(plan.m_r); hi = np.empty(plan.m_r)
    for r, i in enumerate(plan.keep):
        lo[r], hi[r] = row_bounds(s, i, slots_k, days, vol_k)
    lc = np.zeros(n); uc = np.full(n, np.inf)
    bad = False
    for i, j, a in zip(plan.sing_row, plan.sing_col, plan.sing_coef):
        l, u = row_bounds(s, i, slots_k, days, vol_k)
        bad = bad or np.isnan(l) or np.isnan(u)   # max / min below would swallow a NaN (pdhg_assemble_cols: `bad |= isnan`)
        l, u = (l / a, u / a) if a > 0 else (u / a, l / a)
        lc[j] = max(lc[j], l); uc[j] = min(uc[j], u)
    c = np.zeros(n)
    for j in range(n):
        tot = 0.0
        for col, ratio in plan.members[j]:          # the column itself and every column folded into it
            acc = 0.0
            for t in range(s.cost_indptr[col], s.cost_indptr[col + 1]):
                acc += s.cost_w[t] * _ref_value(s, s.cost_ref[t], slots_k)
            if s.cost_clip and abs(acc) < 1e-8:
                acc = 0.0
            tot += ratio * acc
        c[j] = tot
    dead = plan.alias_root != np.arange(n)
    uc[dead] = 0.0                                  # folded columns: fixed at 0 in the reduced LP, rebuilt in snap_unscale
    status = ST_OPTIMAL
    if bad or np.isnan(lo).any() or np.isnan(hi).any() or np.isnan(lc).any() or np.isnan(uc).any() or np.isnan(c).any():
        status = ST_NAN
    elif (lo > hi).any() or (lc > uc + 1e-9).any():
        status = ST_INFEASIBLE
    uc = np.maximum(uc, lc)
    return lo * plan.dr, hi * plan.dr, c * plan.dc, lc / plan.dc, uc / plan.dc, status


def pdhg_operator(plan, x, y, tau, sig, lo, hi, c, lc, uc):
    xn = np.clip(x - tau * (c - plan.AT @ y), lc, uc)
    v = y - sig * (plan.A @ (2.0 * xn - x))
    yn = v - np.clip(v, -sig * hi, -sig * lo)
    return xn, yn


def solve_scaled(plan, lo, hi, c, lc, uc, x, y, omega, valid, tol, max_iter, check):
    """Steps 3-4 for one scenario.  Returns (x, y, omega, iterations, status, restarts)."""
    A = plan.A
    eta = plan.eta
    brow = np.maximum(np.abs(_fin(lo)), np.abs(_fin(hi)))
    bnorm = np.sqrt(brow @ brow + _fin(lc) @ _fin(lc) + _fin(uc) @ _fin(uc))
    cnorm = np.sqrt(c @ c)
    if not valid:
        x = np.zeros(plan.n); y = np.zeros(plan.m_r)
        omega = cnorm / bnorm if (bnorm > 0 and cnorm > 0) else 1.0
    x = np.clip(x, lc, uc)
    xa, ya = x.copy(), y.copy()
    kbase, it, r0, rprev, restarts = 0, 0, -1.0, np.inf, 0
    lo_inf, hi_inf = ~(np.abs(lo) < BIG), ~(np.abs(hi) < BIG)
    lc_inf, uc_inf = ~(np.abs(lc) < BIG), ~(np.abs(uc) < BIG)
    while True:
        tau, sig = eta / omega, eta * omega
        for _ in range(check):
            k = it - kbase
            w = (k + 1.0) / (k + 2.0)
            xn, yn = pdhg_operator(plan, x, y, tau, sig, lo, hi, c, lc, uc)
            x = w * (2.0 * xn - x) + (1.0 - w) * xa
            y = w * (2.0 * yn - y) + (1.0 - w) * ya
            it += 1
        # ---- check at T(z) ----
        xn, yn = pdhg_operator(plan, x, y, tau, sig, lo, hi, c, lc, uc)
        dx, dy = xn - x, yn - y
        axn = A @ xn
        r2 = dx @ dx / tau - 2.0 * (dy @ (axn - A @ x)) + dy @ dy / sig
        r = np.sqrt(max(r2, 0.0))
        pr = np.linalg.norm(axn - np.clip(axn, lo, hi))
        rc = c - plan.AT @ yn
        rp, rm = np.maximum(rc, 0.0), np.maximum(-rc, 0.0)
        yp, ym = np.maximum(yn, 0.0), np.maximum(-yn, 0.0)
        dres2 = rp[lc_inf] @ rp[lc_inf] + rm[uc_inf] @ rm[uc_inf] + yp[lo_inf] @ yp[lo_inf] + ym[hi_inf] @ ym[hi_inf]
        dobj = _fin(lo) @ yp - _fin(hi) @ ym + _fin(lc) @ rp - _fin(uc) @ rm
        pobj = c @ xn
        rel = max(pr / (1.0 + bnorm), np.sqrt(dres2) / (1.0 + cnorm), abs(pobj - dobj) / (1.0 + abs(pobj) + abs(dobj)))
        if r0 < 0.0:
            r0 = r
        k = it - kbase
        if rel < tol or it >= max_iter:
            return xn, yn, omega, it, (ST_OPTIMAL if rel < tol else ST_ITERLIMIT), restarts
        if r <= 0.2 * r0 or (r <= 0.8 * r0 and r > rprev) or k >= 0.36 * it:
            ddx, ddy = np.linalg.norm(xn - xa), np.linalg.norm(yn - ya)
            # update the primal weight only when both iterates really moved since the last restart (a warm-started,
            # already optimal x barely moves and the ratio of two noise terms would blow omega up), bounded around
            # the initial weight
            om0 = cnorm / bnorm if (bnorm > 0 and cnorm > 0) else 1.0
            mx, my = ddx > OMEGA_MOVE * np.linalg.norm(xn), ddy > OMEGA_MOVE * np.linalg.norm(yn)
            e_p, e_d, e_g = pr / (1.0 + bnorm), np.sqrt(dres2) / (1.0 + cnorm), abs(pobj - dobj) / (1.0 + abs(pobj) + abs(dobj))
            if mx and my:
                omega = np.exp(0.5 * np.log(ddy / ddx) + 0.5 * np.log(omega))
            elif my and e_p >= max(e_d, e_g):
                omega *= OMEGA_KICK      # x pinned at its bounds, primal infeasible: let the dual move faster
            elif mx and e_p < max(e_d, e_g):
                omega /= OMEGA_KICK      # y settled, primal feasible but not optimal: let x move faster
            omega = min(max(omega, om0 / OMEGA_SPAN), om0 * OMEGA_SPAN)
            x, y = xn.copy(), yn.copy()
            xa, ya = xn.copy(), yn.copy()
            kbase, r0, restarts = it, -1.0, restarts + 1
        rprev = r


def snap_unscale(plan, xs, lc, uc):
    """Step 5: columns within 1e-9 (relative) of a bound go onto it; un-scale."""
    x = xs * plan.dc
    l, u = lc * plan.dc, uc * plan.dc
    x = np.where(np.abs(x - l) <= 1e-9 * (1.0 + np.abs(l)), l, x)
    x = np.where(np.abs(x - u) <= 1e-9 * (1.0 + np.abs(_fin(u))), u, x)
    out = x.copy()
    for q in np.nonzero(plan.alias_root != np.arange(plan.n))[0]:   # folded columns: x_q = sum rho_k x_k over the surviving columns
        acc = None
        for col, ratio in plan.comb[q]:
            term = ratio * x[col]
            acc = term if acc is None else acc + term
        out[q] = acc
    return out


class PdhgOracleEngine:
    """Same duck-typed interface as ``pywr_b200.engine.CudaEngine`` (per-timestep path)."""

    def __init__(self, structure, n_scenarios, tol=1e-8, max_iter=200000, check=64, **_):
        self.structure = structure
        self.S = int(n_scenarios)
        self.plan = PdhgPlan(structure)
        self.tol, self.max_iter, self.check = tol, max_iter, check
        self.flowmat = sp.csr_matrix((structure.flow_w, structure.flow_col, structure.flow_indptr),
                                     shape=(structure.n_nodes, structure.n_cols))
        self.reset(None)
        self.iterations = 0
        self.last_iterations = np.zeros(self.S, np.int64)

    def alloc_planes(self, n):
        return np.zeros((n, self.S))

    def reset(self, vol0=None):
        p = self.plan
        self.x = np.zeros((self.S, p.n)); self.y = np.zeros((self.S, p.m_r))
        self.omega = np.ones(self.S); self.valid = np.zeros(self.S, bool)
        self.vol = np.zeros((max(self.structure.n_storage, 1), self.S))
        if vol0 is not None:
            self.vol[:] = np.asarray(vol0).reshape(self.vol.shape)
        self._status = np.zeros(self.S, np.int32)

    def set_consts(self, consts):
        self.structure.consts = np.asarray(consts, float).copy()
        self.plan = PdhgPlan(self.structure)
        self.valid[:] = False

    def set_volume(self, vol):
        self.vol[:] = np.asarray(vol).reshape(self.vol.shape)

    def step(self, days, slots, node_flow=None, col_flow=None, objective=None):
        p, s = self.plan, self.structure
        nbad = 0
        for k in range(self.S):
            lo, hi, c, lc, uc, status = assemble(p, slots[:, k], days, self.vol[:, k])
            if status == ST_OPTIMAL:
                xs, ys, om, it, status, _ = solve_scaled(p, lo, hi, c, lc, uc, self.x[k], self.y[k], self.omega[k],
                                                         self.valid[k], self.tol, self.max_iter, self.check)
                self.x[k], self.y[k], self.omega[k], self.valid[k] = xs, ys, om, True
                self.iterations += it
                self.last_iterations[k] = it
            self._status[k] = status
            if status != ST_OPTIMAL:
                nbad += 1
                continue
            x = snap_unscale(p, self.x[k], lc, uc)
            if col_flow is not None:
                col_flow[:, k] = x
            if node_flow is not None:
                node_flow[:, k] = self.flowmat @ x
            if objective is not None:
                objective[0, k] = (c / p.dc) @ x
        return nbad

    def status(self):
        bad = np.nonzero(self._status != 0)[0]
        return (int(bad[0]), int(self._status[bad[0]])) if len(bad) else (-1, 0)

    def stats(self):
        return {"pdhg_iterations": int(self.iterations), "n_rows": self.structure.n_rows, "n_cols": self.structure.n_cols,
                "reduced_rows": self.plan.m_r}

    def close(self):
        pass
