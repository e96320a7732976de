/*
 * lp_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Scalar C restatement of the hot path of pywr/pywr that this repository replaces on the GPU:
 * the per-timestep, per-scenario loop of CythonGLPKSolver.solve.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load this library; the
 * product (pywr_b200/) never does.
 *
 * What it follows in the reference (paths relative to /root/reference):
 *   - bound typing / clipping            pywr/solvers/cython_glpk.pyx:66-95,932-959   (a5)
 *   - storage / virtual storage bounds   pywr/solvers/cython_glpk.pyx:965-1019        (a6)
 *   - objective assembly                 pywr/solvers/cython_glpk.pyx:880-897         (a3)
 *   - aggregated-factor matrix rewrite   pywr/solvers/cython_glpk.pyx:903-926         (a4)
 *   - basis carry-over per scenario      pywr/solvers/cython_glpk.pyx:197-232,1026,1063 (a7)
 *   - result extraction                  pywr/solvers/cython_glpk.pyx:1068-1093       (a8)
 *   - edge formulation equivalents       pywr/solvers/cython_glpk.pyx:1649-1918
 *
 * The LP arithmetic itself lives in GNU GLPK (glp_simplex), a third-party C library that is NOT
 * vendored in the reference and is NOT installed in this image (no glpk.h / libglpk; the reference
 * does not pin a version: setup.py:76-83 `libraries=["glpk"]`, CI uses glpk 5.0).  GLPK's simplex
 * is therefore restated here by its published contract — minimise c'x s.t. lb <= Ax <= ub, x >= 0,
 * warm-started from the scenario's previous basis, optimal vertex returned — as a dense bounded
 * dual simplex with a primal clean-up phase.  Parity is anchored on the reference's own
 * known-answer tests (tests/test_analytical.py, test_run.py, test_scenarios.py,
 * test_agg_constraints.py, ... run with this oracle plugged into the built reference through
 * oracle/pywr_oracle_solver.py) and cross-checked against scipy's HiGHS.
 * Parity with GLPK on the long benchmark configurations is UNPINNED (GLPK cannot run here).
 *
 * Dictionary form.  Variables v = (x_0..x_{n-1}, r_0..r_{m-1}) with r = A x.  A basis is a split
 * into m basic and n non-basic variables; T (m x n, dense) expresses the basic variables through
 * the non-basic ones, v_head[i] = sum_j T[i][j] v_nb[j].  The slack basis is T = A.  All arithmetic
 * that the CUDA kernels mirror is written with explicit fma() in a fixed order so that both sides
 * can agree to the last bit.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200lp.h"

#define ORC_TOL_PRIMAL 1e-9
#define ORC_TOL_DUAL 1e-9
#define ORC_TOL_PIVOT 1e-9
#define ORC_TIE_REL 1e-9
#define ORC_TIE_ABS 1e-12
#define ORC_REFACTOR_PIVOTS 40

enum { VS_BASIC = 1, VS_NL = 2, VS_NU = 3, VS_NF = 4, VS_NS = 5 }; /* GLPK's GLP_BS.. codes */

typedef struct {
    double *T;     /* [m*n] dictionary                               */
    double *d;     /* [n] reduced costs for the true objective       */
    int32_t *head; /* [m] variable in basic position i               */
    int32_t *nb;   /* [n] variable in non-basic position j           */
    int8_t *vstat; /* [n+m] per-variable status (the "basis")        */
    double *vol;   /* [n_storage] device-state volume equivalent     */
    double *pc;    /* [n_storage] volume / max_volume (Storage._current_pc) */
    int32_t npiv;  /* pivots since the last re-factorisation         */
    int32_t dead;  /* program mode: failed at an earlier timestep           */
    int32_t valid; /* T/d/head/nb consistent with vstat              */
    int32_t status;
} orc_scen;

typedef struct {
    b200lp_structure s; /* deep copy */
    int32_t S, m, n, NV, nnz;
    double *A; /* dense m*n image of a_val (static part) */
    orc_scen *sc;
    int64_t pivots, refactors, steps;
    int32_t first_bad, bad_code;
    int32_t threads;
} orc_handle;

/* ------------------------------------------------------------------ helpers */

static void *dup_mem(const void *p, size_t bytes) {
    void *q = malloc(bytes ? bytes : 1);
    if (p && bytes) memcpy(q, p, bytes);
    return q;
}

/* slots: planes [n_slots][S] (stride S, offset = scenario) in step mode, or this scenario's own
 * slot vector (stride 1, offset 0) in program mode */
typedef struct { const double *base; size_t stride, off; } slot_view;

static inline double ref_value_v(const orc_handle *h, int32_t ref, slot_view sv) {
    if (ref >= 0) return sv.base[(size_t)ref * sv.stride + sv.off];
    return h->s.consts[-ref - 1];
}
#define ref_value(h, ref, slots, sidx) ref_value_v(h, ref, sv)

/* cython_glpk.pyx:934-945 + constraint_type :66-77: clip tiny values, FX when |lb-ub| < 1e-8 */
static inline void type_bounds(double *lo, double *hi) {
    if (fabs(*lo) < 1e-8) *lo = 0.0;
    if (fabs(*hi) < 1e-8) *hi = 0.0;
    if (fabs(*lo - *hi) < 1e-8) *hi = *lo;
}

/* ------------------------------------------------------------------ create / destroy */

void *orc_create(const b200lp_structure *s, int32_t S) {
    if (!s || s->abi_version != B200LP_ABI_VERSION || S <= 0) return NULL;
    orc_handle *h = (orc_handle *)calloc(1, sizeof(orc_handle));
    h->s = *s;
    int32_t m = s->n_rows, n = s->n_cols, N = s->n_nodes;
    h->S = S; h->m = m; h->n = n; h->NV = m + n;
    h->nnz = s->a_indptr[m];
    h->s.a_indptr = dup_mem(s->a_indptr, sizeof(int32_t) * (m + 1));
    h->s.a_col = dup_mem(s->a_col, sizeof(int32_t) * h->nnz);
    h->s.a_val = dup_mem(s->a_val, sizeof(double) * h->nnz);
    h->s.row_kind = dup_mem(s->row_kind, sizeof(int32_t) * m);
    h->s.row_lb_ref = dup_mem(s->row_lb_ref, sizeof(int32_t) * m);
    h->s.row_ub_ref = dup_mem(s->row_ub_ref, sizeof(int32_t) * m);
    h->s.st_kind = dup_mem(s->st_kind, sizeof(int32_t) * s->n_storage);
    h->s.st_vol_ref = dup_mem(s->st_vol_ref, sizeof(int32_t) * s->n_storage);
    h->s.st_minv_ref = dup_mem(s->st_minv_ref, sizeof(int32_t) * s->n_storage);
    h->s.st_maxv_ref = dup_mem(s->st_maxv_ref, sizeof(int32_t) * s->n_storage);
    h->s.st_active_ref = dup_mem(s->st_active_ref, sizeof(int32_t) * s->n_storage);
    int32_t ncost = s->cost_indptr[n];
    h->s.cost_indptr = dup_mem(s->cost_indptr, sizeof(int32_t) * (n + 1));
    h->s.cost_ref = dup_mem(s->cost_ref, sizeof(int32_t) * ncost);
    h->s.cost_w = dup_mem(s->cost_w, sizeof(double) * ncost);
    h->s.dyn_a_nz = dup_mem(s->dyn_a_nz, sizeof(int32_t) * s->n_dyn_a);
    h->s.dyn_a_ref = dup_mem(s->dyn_a_ref, sizeof(int32_t) * s->n_dyn_a);
    h->s.dyn_a_scale = dup_mem(s->dyn_a_scale, sizeof(double) * s->n_dyn_a);
    int32_t nflow = s->flow_indptr[N];
    h->s.flow_indptr = dup_mem(s->flow_indptr, sizeof(int32_t) * (N + 1));
    h->s.flow_col = dup_mem(s->flow_col, sizeof(int32_t) * nflow);
    h->s.flow_w = dup_mem(s->flow_w, sizeof(double) * nflow);
    h->s.consts = dup_mem(s->consts, sizeof(double) * s->n_consts);

    h->A = (double *)calloc((size_t)m * n + 1, sizeof(double));
    for (int32_t i = 0; i < m; i++)
        for (int32_t k = s->a_indptr[i]; k < s->a_indptr[i + 1]; k++)
            h->A[(size_t)i * n + s->a_col[k]] += s->a_val[k];

    h->sc = (orc_scen *)calloc(S, sizeof(orc_scen));
    for (int32_t k = 0; k < S; k++) {
        orc_scen *c = &h->sc[k];
        c->T = (double *)calloc((size_t)m * n + 1, sizeof(double));
        c->d = (double *)calloc(n + 1, sizeof(double));
        c->head = (int32_t *)calloc(m + 1, sizeof(int32_t));
        c->nb = (int32_t *)calloc(n + 1, sizeof(int32_t));
        c->vstat = (int8_t *)calloc(m + n + 1, sizeof(int8_t));
        c->vol = (double *)calloc(s->n_storage + 1, sizeof(double));
        c->pc = (double *)calloc(s->n_storage + 1, sizeof(double));
    }
    h->first_bad = -1;
    return h;
}

void orc_unload_program(void *hv);

void orc_destroy(void *hv) {
    orc_handle *h = (orc_handle *)hv;
    if (!h) return;
    orc_unload_program(hv);
    for (int32_t k = 0; k < h->S; k++) {
        orc_scen *c = &h->sc[k];
        free(c->T); free(c->d); free(c->head); free(c->nb); free(c->vstat); free(c->vol); free(c->pc);
    }
    free(h->sc); free(h->A);
    free((void *)h->s.a_indptr); free((void *)h->s.a_col); free((void *)h->s.a_val);
    free((void *)h->s.row_kind); free((void *)h->s.row_lb_ref); free((void *)h->s.row_ub_ref);
    free((void *)h->s.st_kind); free((void *)h->s.st_vol_ref); free((void *)h->s.st_minv_ref);
    free((void *)h->s.st_maxv_ref); free((void *)h->s.st_active_ref);
    free((void *)h->s.cost_indptr); free((void *)h->s.cost_ref); free((void *)h->s.cost_w);
    free((void *)h->s.dyn_a_nz); free((void *)h->s.dyn_a_ref); free((void *)h->s.dyn_a_scale);
    free((void *)h->s.flow_indptr); free((void *)h->s.flow_col); free((void *)h->s.flow_w);
    free((void *)h->s.consts);
    free(h);
}

/* Solver.reset(): is_first_solve = True (cython_glpk.pyx:757-759) -> every scenario restarts from
 * the slack basis (all rows basic, all columns at their lower bound 0). */
void orc_reset(void *hv, const double *vol0) {
    orc_handle *h = (orc_handle *)hv;
    for (int32_t k = 0; k < h->S; k++) {
        orc_scen *c = &h->sc[k];
        for (int32_t j = 0; j < h->n; j++) c->vstat[j] = VS_NL;
        for (int32_t i = 0; i < h->m; i++) c->vstat[h->n + i] = VS_BASIC;
        c->valid = 0; c->npiv = 0; c->status = 0;
        for (int32_t q = 0; q < h->s.n_storage; q++) c->vol[q] = vol0 ? vol0[(size_t)q * h->S + k] : 0.0;
    }
    h->first_bad = -1; h->bad_code = 0;
}

/* ------------------------------------------------------------------ the dictionary simplex */

typedef struct {
    int32_t m, n;
    double *T, *d, *dw;
    int32_t *head, *nb;
    int8_t *vstat;
    const double *lo, *hi; /* [NV] bounds per variable */
    double *xn;            /* [n] value of each non-basic position */
    double *beta;          /* [m] value of each basic position     */
    int64_t pivots;
} dict;

/* exchange basic position r with non-basic position q */
static void dict_pivot(dict *D, int32_t r, int32_t q, int use_dw) {
    const int32_t m = D->m, n = D->n;
    double *T = D->T;
    double *prow = T + (size_t)r * n;
    const double inv = 1.0 / prow[q];
    for (int32_t j = 0; j < n; j++)
        if (j != q) prow[j] = -prow[j] * inv;
    prow[q] = inv;
    for (int32_t i = 0; i < m; i++) {
        if (i == r) continue;
        double *row = T + (size_t)i * n;
        const double f = row[q];
        for (int32_t j = 0; j < n; j++)
            if (j != q) row[j] = fma(f, prow[j], row[j]);
        row[q] = f * inv;
    }
    {
        const double f = D->d[q];
        for (int32_t j = 0; j < n; j++)
            if (j != q) D->d[j] = fma(f, prow[j], D->d[j]);
        D->d[q] = f * inv;
    }
    if (use_dw) {
        const double f = D->dw[q];
        for (int32_t j = 0; j < n; j++)
            if (j != q) D->dw[j] = fma(f, prow[j], D->dw[j]);
        D->dw[q] = f * inv;
    }
    int32_t t = D->head[r]; D->head[r] = D->nb[q]; D->nb[q] = t;
    D->pivots++;
}

/* d_j = c_nb[j] + sum_i c_head[i] T[i][j]  (row variables have zero cost) */
static void dict_reduced_costs(dict *D, const double *c) {
    const int32_t m = D->m, n = D->n;
    for (int32_t j = 0; j < n; j++) {
        double acc = (D->nb[j] < n) ? c[D->nb[j]] : 0.0;
        for (int32_t i = 0; i < m; i++) {
            const int32_t v = D->head[i];
            if (v < n) acc = fma(c[v], D->T[(size_t)i * n + j], acc);
        }
        D->d[j] = acc;
    }
}

/* Rebuild T for the basis recorded in vstat by pivoting from the slack basis (Gauss-Jordan with
 * partial pivoting over the rows that must become non-basic).  Returns 0, or -1 when singular. */
static int dict_refactor(dict *D, const double *A) {
    const int32_t m = D->m, n = D->n;
    memcpy(D->T, A, sizeof(double) * (size_t)m * n);
    for (int32_t i = 0; i < m; i++) D->head[i] = n + i;
    for (int32_t j = 0; j < n; j++) D->nb[j] = j;
    for (int32_t j = 0; j < n; j++) D->d[j] = 0.0; /* recomputed by the caller */
    for (int32_t q = 0; q < n; q++) {
        if (D->vstat[q] != VS_BASIC) continue;
        int32_t r = -1; double best = ORC_TOL_PIVOT;
        for (int32_t i = 0; i < m; i++) {
            const int32_t v = D->head[i];
            if (v < n || D->vstat[v] == VS_BASIC) continue; /* slot already used / row stays basic */
            const double a = fabs(D->T[(size_t)i * n + q]);
            if (a > best) { best = a; r = i; }
        }
        if (r < 0) return -1;
        dict_pivot(D, r, q, 0);
        D->pivots--; /* not a simplex pivot */
    }
    return 0;
}

static void dict_slack_basis(dict *D) {
    for (int32_t j = 0; j < D->n; j++) D->vstat[j] = VS_NL;
    for (int32_t i = 0; i < D->m; i++) D->vstat[D->n + i] = VS_BASIC;
}

/* status + value of every non-basic position for the current bounds; returns 1 if some reduced
 * cost in `dj` is dual infeasible after the boxed variables have been put on their better bound */
static int dict_place_nonbasics(dict *D, const double *dj) {
    int infeas = 0;
    for (int32_t j = 0; j < D->n; j++) {
        const int32_t v = D->nb[j];
        const double lo = D->lo[v], hi = D->hi[v];
        int8_t st = D->vstat[v];
        if (lo == hi) st = VS_NS;
        else if (lo > -INFINITY && hi < INFINITY) {
            if (dj[j] > ORC_TOL_DUAL) st = VS_NL;
            else if (dj[j] < -ORC_TOL_DUAL) st = VS_NU;
            else if (st != VS_NL && st != VS_NU) st = VS_NL;
        } else if (lo > -INFINITY) { st = VS_NL; if (dj[j] < -ORC_TOL_DUAL) infeas = 1; }
        else if (hi < INFINITY) { st = VS_NU; if (dj[j] > ORC_TOL_DUAL) infeas = 1; }
        else { st = VS_NF; if (fabs(dj[j]) > ORC_TOL_DUAL) infeas = 1; }
        D->vstat[v] = st;
        D->xn[j] = (st == VS_NU) ? hi : (st == VS_NF ? 0.0 : lo);
    }
    return infeas;
}

static void dict_basic_values(dict *D) {
    const int32_t m = D->m, n = D->n;
    for (int32_t i = 0; i < m; i++) {
        double acc = 0.0;
        const double *row = D->T + (size_t)i * n;
        for (int32_t j = 0; j < n; j++) acc = fma(row[j], D->xn[j], acc);
        D->beta[i] = acc;
    }
}

/* Runs the simplex from the current dictionary.  Returns a B200LP_ST_* code. */
static int dict_solve(dict *D) {
    const int32_t m = D->m, n = D->n;
    const int32_t cap_bland = 20 * (m + n) + 50, cap_fail = 200 * (m + n) + 1000;
    int shifted = dict_place_nonbasics(D, D->d);
    /* cost shifting: make the working reduced costs dual feasible, fix the rest in the primal phase */
    for (int32_t j = 0; j < n; j++) {
        double w = D->d[j];
        if (shifted) {
            const int8_t st = D->vstat[D->nb[j]];
            if ((st == VS_NL && w < -ORC_TOL_DUAL) || (st == VS_NU && w > ORC_TOL_DUAL) ||
                (st == VS_NF && fabs(w) > ORC_TOL_DUAL)) w = 0.0;
        }
        D->dw[j] = w;
    }
    int phase = 0; /* 0: dual simplex on dw, 1: primal simplex on d */
    for (int32_t it = 0;; it++) {
        if (it > cap_fail) return B200LP_ST_ITERLIMIT;
        const int bland = it > cap_bland;
        dict_basic_values(D);
        if (phase == 0) {
            /* leaving variable: largest bound violation */
            int32_t r = -1; double worst = 0.0; int dir = 0;
            for (int32_t i = 0; i < m; i++) {
                const int32_t v = D->head[i];
                const double b = D->beta[i];
                double viol = 0.0; int dd = 0;
                if (b < D->lo[v] - ORC_TOL_PRIMAL * (1.0 + fabs(D->lo[v]))) { viol = D->lo[v] - b; dd = +1; }
                else if (b > D->hi[v] + ORC_TOL_PRIMAL * (1.0 + fabs(D->hi[v]))) { viol = b - D->hi[v]; dd = -1; }
                if (dd != 0 && (bland ? (r < 0) : (viol > worst))) { worst = viol; r = i; dir = dd; }
            }
            if (r < 0) {
                if (!shifted) return B200LP_ST_OPTIMAL;
                /* primal feasible for the shifted costs: finish with the primal simplex on the
                 * true reduced costs (non-basic variables keep their bounds: moving a boxed one
                 * now would destroy primal feasibility; it enters and flips in the ratio test) */
                phase = 1; shifted = 0;
                continue;
            }
            /* dual ratio test on row r */
            const double *prow = D->T + (size_t)r * n;
            double rmin = INFINITY;
            for (int32_t j = 0; j < n; j++) {
                const int8_t st = D->vstat[D->nb[j]];
                const double a = prow[j] * (double)dir;
                int ok = 0;
                if (st == VS_NL) ok = a > ORC_TOL_PIVOT;
                else if (st == VS_NU) ok = a < -ORC_TOL_PIVOT;
                else if (st == VS_NF) ok = fabs(a) > ORC_TOL_PIVOT;
                if (!ok) continue;
                const double ratio = fabs(D->dw[j]) / fabs(a);
                if (ratio < rmin) rmin = ratio;
            }
            if (rmin == INFINITY) return B200LP_ST_INFEASIBLE;
            const double thr = rmin * (1.0 + ORC_TIE_REL) + ORC_TIE_ABS;
            int32_t q = -1; double amax = 0.0;
            for (int32_t j = 0; j < n; j++) {
                const int8_t st = D->vstat[D->nb[j]];
                const double a = prow[j] * (double)dir;
                int ok = 0;
                if (st == VS_NL) ok = a > ORC_TOL_PIVOT;
                else if (st == VS_NU) ok = a < -ORC_TOL_PIVOT;
                else if (st == VS_NF) ok = fabs(a) > ORC_TOL_PIVOT;
                if (!ok) continue;
                const double ratio = fabs(D->dw[j]) / fabs(a);
                if (ratio > thr) continue;
                if (bland ? (q < 0) : (fabs(a) > amax)) { amax = fabs(a); q = j; }
            }
            const int32_t vout = D->head[r];
            dict_pivot(D, r, q, 1);
            D->vstat[D->head[r]] = VS_BASIC;
            D->vstat[vout] = (D->lo[vout] == D->hi[vout]) ? VS_NS : (dir > 0 ? VS_NL : VS_NU);
            D->xn[q] = dir > 0 ? D->lo[vout] : D->hi[vout];
        } else {
            /* entering variable: largest dual infeasibility (Dantzig) */
            int32_t q = -1; double worst = 0.0; int dir = 0;
            for (int32_t j = 0; j < n; j++) {
                const int8_t st = D->vstat[D->nb[j]];
                const double dj = D->d[j];
                int dd = 0;
                if ((st == VS_NL || st == VS_NF) && dj < -ORC_TOL_DUAL) dd = +1;
                else if ((st == VS_NU || st == VS_NF) && dj > ORC_TOL_DUAL) dd = -1;
                if (dd != 0 && (bland ? (q < 0) : (fabs(dj) > worst))) { worst = fabs(dj); q = j; dir = dd; }
            }
            if (q < 0) return B200LP_ST_OPTIMAL;
            const int32_t vin = D->nb[q];
            /* primal ratio test down column q */
            double tmin = D->hi[vin] - D->lo[vin]; /* own range: bound flip */
            if (!(tmin < INFINITY)) tmin = INFINITY;
            double step_min = INFINITY;
            for (int32_t i = 0; i < m; i++) {
                const int32_t v = D->head[i];
                const double a = D->T[(size_t)i * n + q] * (double)dir;
                double step = INFINITY;
                if (a > ORC_TOL_PIVOT) { if (D->hi[v] < INFINITY) step = (D->hi[v] - D->beta[i]) / a; }
                else if (a < -ORC_TOL_PIVOT) { if (D->lo[v] > -INFINITY) step = (D->lo[v] - D->beta[i]) / a; }
                if (step < 0.0) step = 0.0;
                if (step < step_min) step_min = step;
            }
            if (step_min == INFINITY && tmin == INFINITY) return B200LP_ST_UNBOUNDED;
            if (tmin <= step_min) {
                /* the entering variable reaches its own opposite bound first: no basis change */
                D->vstat[vin] = dir > 0 ? VS_NU : VS_NL;
                D->xn[q] = dir > 0 ? D->hi[vin] : D->lo[vin];
                continue;
            }
            const double thr = step_min * (1.0 + ORC_TIE_REL) + ORC_TIE_ABS;
            int32_t r = -1; double amax = 0.0; int hit_upper = 0;
            for (int32_t i = 0; i < m; i++) {
                const int32_t v = D->head[i];
                const double a = D->T[(size_t)i * n + q] * (double)dir;
                double step = INFINITY; int up = 0;
                if (a > ORC_TOL_PIVOT) { if (D->hi[v] < INFINITY) { step = (D->hi[v] - D->beta[i]) / a; up = 1; } }
                else if (a < -ORC_TOL_PIVOT) { if (D->lo[v] > -INFINITY) step = (D->lo[v] - D->beta[i]) / a; }
                if (step < 0.0) step = 0.0;
                if (step > thr) continue;
                if (bland ? (r < 0) : (fabs(a) > amax)) { amax = fabs(a); r = i; hit_upper = up; }
            }
            const int32_t vout = D->head[r];
            dict_pivot(D, r, q, 0);
            D->vstat[D->head[r]] = VS_BASIC;
            D->vstat[vout] = (D->lo[vout] == D->hi[vout]) ? VS_NS : (hit_upper ? VS_NU : VS_NL);
            D->xn[q] = hit_upper ? D->hi[vout] : D->lo[vout];
        }
    }
}

/* ------------------------------------------------------------------ one scenario, one timestep */

static int orc_solve_scenario(orc_handle *h, int32_t sidx, double days, slot_view sv, double *node_flow,
                              double *col_flow, double *objective, double *work, int64_t *pivots,
                              int64_t *refactors) {
    const b200lp_structure *s = &h->s;
    const int32_t m = h->m, n = h->n, NV = h->NV, N = s->n_nodes, S = h->S;
    orc_scen *c = &h->sc[sidx];
    double *lo = work, *hi = lo + NV, *cost = hi + NV, *xn = cost + n, *beta = xn + n, *dw = beta + m,
           *x = dw + n, *ract = x + n, *Ad = ract + m;
    int nan_seen = 0, bad_bounds = 0;

    /* (a5/a6) bounds: columns are x >= 0 (cython_glpk.pyx:490-491) */
    for (int32_t j = 0; j < n; j++) { lo[j] = 0.0; hi[j] = INFINITY; }
    for (int32_t i = 0; i < m; i++) {
        double l, u;
        if (s->row_kind[i] == B200LP_ROW_FLOW) {
            l = ref_value(h, s->row_lb_ref[i], slots, sidx);
            u = ref_value(h, s->row_ub_ref[i], slots, sidx);
            if (isnan(l) || isnan(u)) nan_seen = 1;
            type_bounds(&l, &u);
        } else {
            const int32_t k = s->row_lb_ref[i];
            const double maxv = ref_value(h, s->st_maxv_ref[k], slots, sidx);
            const double minv = ref_value(h, s->st_minv_ref[k], slots, sidx);
            const double vol = (s->st_vol_ref[k] == B200LP_REF_STATE) ? c->vol[k]
                                                                       : ref_value(h, s->st_vol_ref[k], slots, sidx);
            int active = 1;
            if (s->st_kind[k] == B200LP_ST_KIND_VIRTUAL) active = ref_value(h, s->st_active_ref[k], slots, sidx) != 0.0;
            if (isnan(maxv) || isnan(minv) || isnan(vol)) nan_seen = 1;
            if (maxv == minv) { l = 0.0; u = 0.0; }
            else if (!active) { l = -INFINITY; u = INFINITY; }
            else {
                const double avail = fmax(vol - minv, 0.0);
                l = -avail / days;
                u = fmax(maxv - vol, 0.0) / days;
                type_bounds(&l, &u);
            }
        }
        if (l > u) bad_bounds = 1;
        lo[n + i] = l; hi[n + i] = u;
    }
    /* (a3) objective */
    int cost_dynamic = 0;
    for (int32_t j = 0; j < n; j++) {
        double acc = 0.0;
        for (int32_t k = s->cost_indptr[j]; k < s->cost_indptr[j + 1]; k++) {
            if (s->cost_ref[k] >= 0) cost_dynamic = 1;
            acc += s->cost_w[k] * ref_value(h, s->cost_ref[k], slots, sidx);
        }
        if (isnan(acc)) nan_seen = 1;
        if (s->cost_clip && fabs(acc) < 1e-8) acc = 0.0;
        cost[j] = acc;
    }
    /* (a4) dynamic matrix entries */
    const double *A = h->A;
    if (s->n_dyn_a > 0) {
        memcpy(Ad, h->A, sizeof(double) * (size_t)m * n);
        for (int32_t k = 0; k < s->n_dyn_a; k++) {
            const int32_t nz = s->dyn_a_nz[k];
            int32_t row = 0;
            while (s->a_indptr[row + 1] <= nz) row++;
            const double v = s->dyn_a_scale[k] * ref_value(h, s->dyn_a_ref[k], slots, sidx);
            if (isnan(v)) nan_seen = 1;
            /* replace the static contribution of this non-zero */
            Ad[(size_t)row * n + s->a_col[nz]] += v - s->a_val[nz];
        }
        A = Ad;
    }
    if (nan_seen) { c->status = B200LP_ST_NAN; return c->status; }
    if (bad_bounds) { c->status = B200LP_ST_INFEASIBLE; return c->status; }

    dict D = {m, n, c->T, c->d, dw, c->head, c->nb, c->vstat, lo, hi, xn, beta, 0};
    int need_refactor = !c->valid || s->n_dyn_a > 0 || c->npiv >= ORC_REFACTOR_PIVOTS;
    const int fresh0 = need_refactor;
    int st = 0;
    /* attempt 0: warm start.  On failure, the role of retry_solve (cython_glpk.pyx:1034-1042):
     * attempt 1 re-factorises the same basis (protects against drift of T), attempt 2 restarts
     * from the standard (slack) basis. */
    for (int attempt = 0; attempt < 3; attempt++) {
        if (attempt == 1 && fresh0) continue;
        if (attempt == 2) dict_slack_basis(&D);
        if (attempt > 0) need_refactor = 1;
        if (need_refactor) {
            if (dict_refactor(&D, A) != 0) {
                dict_slack_basis(&D);
                dict_refactor(&D, A);
            }
            dict_reduced_costs(&D, cost);
            c->npiv = 0; c->valid = 1; (*refactors)++;
        } else if (cost_dynamic) {
            dict_reduced_costs(&D, cost);
        }
        const int64_t p0 = D.pivots;
        st = dict_solve(&D);
        c->npiv += (int32_t)(D.pivots - p0);
        if (st == B200LP_ST_OPTIMAL) break;
    }
    *pivots += D.pivots;
    c->status = st;
    if (st != B200LP_ST_OPTIMAL) { c->valid = 0; return st; }

    /* (a8) primal values of the columns, objective, node flows */
    for (int32_t j = 0; j < n; j++) {
        const int32_t v = c->nb[j];
        if (v < n) x[v] = xn[j];
    }
    dict_basic_values(&D);
    for (int32_t i = 0; i < m; i++) {
        const int32_t v = c->head[i];
        if (v < n) x[v] = beta[i]; else ract[v - n] = beta[i];
    }
    /* activities of the non-basic rows: they sit on a bound */
    for (int32_t j = 0; j < n; j++) {
        const int32_t v = c->nb[j];
        if (v >= n) ract[v - n] = xn[j];
    }
    /* x stays in `work` (ORC_WORK_X) for the program mode */
    if (col_flow) for (int32_t j = 0; j < n; j++) col_flow[(size_t)j * S + sidx] = x[j];
    if (objective) {
        double acc = 0.0;
        for (int32_t j = 0; j < n; j++) acc = fma(cost[j], x[j], acc);
        objective[sidx] = acc;
    }
    if (node_flow) {
        for (int32_t v = 0; v < N; v++) {
            double acc = 0.0;
            for (int32_t k = s->flow_indptr[v]; k < s->flow_indptr[v + 1]; k++)
                acc = fma(s->flow_w[k], x[s->flow_col[k]], acc);
            node_flow[(size_t)v * S + sidx] = acc;
        }
    }
    return st;
}

/* The loop of CythonGLPKSolver.solve (cython_glpk.pyx:821-831), optionally split over host threads
 * (contiguous scenario blocks, the reference-native way to use several cores: Scenario.slice,
 * pywr/_core.pyx:108-128).  Returns the number of failed scenarios. */
typedef struct {
    orc_handle *h;
    int32_t s0, s1;
    double days;
    const double *slots;
    double *node_flow, *col_flow, *objective;
    int nbad;
    int32_t first_bad, bad_code;
    int64_t pivots, refactors;
} orc_job;

static void *orc_worker(void *arg) {
    orc_job *J = (orc_job *)arg;
    orc_handle *h = J->h;
    const int32_t m = h->m, n = h->n;
    double *work = (double *)malloc(sizeof(double) * (2 * (size_t)(m + n) + 4 * n + 2 * m + (size_t)m * n + 8));
    J->nbad = 0; J->first_bad = -1; J->bad_code = 0; J->pivots = 0; J->refactors = 0;
    for (int32_t k = J->s0; k < J->s1; k++) {
        slot_view sv = {J->slots, (size_t)h->S, (size_t)k};
        int st = orc_solve_scenario(h, k, J->days, sv, J->node_flow, J->col_flow, J->objective, work,
                                    &J->pivots, &J->refactors);
        if (st != B200LP_ST_OPTIMAL) {
            J->nbad++;
            if (J->first_bad < 0) { J->first_bad = k; J->bad_code = st; }
        }
    }
    free(work);
    return NULL;
}

int orc_step(void *hv, double days, const double *slots, double *node_flow, double *col_flow, double *objective) {
    orc_handle *h = (orc_handle *)hv;
    int nt = h->threads < 1 ? 1 : h->threads;
    if (nt > h->S) nt = h->S;
    orc_job *jobs = (orc_job *)calloc(nt, sizeof(orc_job));
    pthread_t *tid = (pthread_t *)calloc(nt, sizeof(pthread_t));
    for (int t = 0; t < nt; t++) {
        jobs[t].h = h; jobs[t].days = days; jobs[t].slots = slots;
        jobs[t].node_flow = node_flow; jobs[t].col_flow = col_flow; jobs[t].objective = objective;
        jobs[t].s0 = (int32_t)((int64_t)h->S * t / nt);
        jobs[t].s1 = (int32_t)((int64_t)h->S * (t + 1) / nt);
    }
    for (int t = 1; t < nt; t++) pthread_create(&tid[t], NULL, orc_worker, &jobs[t]);
    orc_worker(&jobs[0]);
    for (int t = 1; t < nt; t++) pthread_join(tid[t], NULL);
    int nbad = 0;
    h->first_bad = -1; h->bad_code = 0;
    for (int t = 0; t < nt; t++) {
        nbad += jobs[t].nbad;
        h->pivots += jobs[t].pivots; h->refactors += jobs[t].refactors;
        if (jobs[t].first_bad >= 0 && h->first_bad < 0) { h->first_bad = jobs[t].first_bad; h->bad_code = jobs[t].bad_code; }
    }
    h->steps++;
    free(jobs); free(tid);
    return nbad;
}

void orc_set_threads(void *hv, int32_t n) { ((orc_handle *)hv)->threads = n; }

void orc_set_consts(void *hv, const double *consts) {
    orc_handle *h = (orc_handle *)hv;
    memcpy((void *)h->s.consts, consts, sizeof(double) * h->s.n_consts);
}

void orc_status(void *hv, int32_t *first_bad, int32_t *code) {
    orc_handle *h = (orc_handle *)hv;
    if (first_bad) *first_bad = h->first_bad;
    if (code) *code = h->bad_code;
}

void orc_get_basis(void *hv, int32_t *row_stat, int32_t *col_stat) {
    orc_handle *h = (orc_handle *)hv;
    for (int32_t k = 0; k < h->S; k++) {
        for (int32_t j = 0; j < h->n; j++) col_stat[(size_t)k * h->n + j] = h->sc[k].vstat[j];
        for (int32_t i = 0; i < h->m; i++) row_stat[(size_t)k * h->m + i] = h->sc[k].vstat[h->n + i];
    }
}

void orc_get_volume(void *hv, double *vol) {
    orc_handle *h = (orc_handle *)hv;
    for (int32_t q = 0; q < h->s.n_storage; q++)
        for (int32_t k = 0; k < h->S; k++) vol[(size_t)q * h->S + k] = h->sc[k].vol[q];
}

void orc_set_volume(void *hv, const double *vol) {
    orc_handle *h = (orc_handle *)hv;
    for (int32_t q = 0; q < h->s.n_storage; q++)
        for (int32_t k = 0; k < h->S; k++) h->sc[k].vol[q] = vol[(size_t)q * h->S + k];
}

void orc_counters(void *hv, int64_t *pivots, int64_t *refactors, int64_t *steps) {
    orc_handle *h = (orc_handle *)hv;
    if (pivots) *pivots = h->pivots;
    if (refactors) *refactors = h->refactors;
    if (steps) *steps = h->steps;
}

/* ================================================================================================
 * Program mode: the whole run per scenario, restating what Pywr does around the solver each
 * timestep (pywr/_model.pyx:626-631,839-860):
 *   before(): parameters are evaluated from the state left by the previous timestep
 *             (ops: pywr/parameters/_parameters.pyx:192-335,1262-1312,1439-1557,
 *              _control_curves.pyx:345-364,501-524; Storage.get_current_pc pywr/_core.pyx:1027-1039)
 *   solve():  orc_solve_scenario
 *   after():  Storage.after mass balance with 1e-6 snapping (pywr/_core.pyx:1335-1361),
 *             VirtualStorage.after (:1483-1493), then the recorders sample
 *             (pywr/recorders/_recorders.pyx:611-617,680-688,723-734,769-779,1175-1183,
 *              1744-1753,1768-1775,1795-1801,1808-1822,1845-1852).
 * ============================================================================================== */
typedef struct {
    b200lp_program p; /* deep copy */
    int loaded;
    double **rec_out; /* [n_recorders] host outputs */
    double *last_x;   /* [n][S] */
    int32_t n_failed, fail_scenario, fail_code, fail_step;
    int32_t n_roll_storage;
    double **roll_mem;   /* [n_storage] ring buffers [len][S] of the rolling virtual storages (NULL: not rolling) */
    int32_t *roll_len;   /* [n_storage] */
    double *roll_init;   /* [n_storage] */
    pthread_mutex_t lock;
} orc_program;

static orc_program *orc_prog_of(orc_handle *h);

/* the handle keeps the program in a side table (keeps orc_handle's layout untouched above) */
#define ORC_MAX_PROGRAMS 64
static struct { orc_handle *h; orc_program *pr; } g_programs[ORC_MAX_PROGRAMS];
static pthread_mutex_t g_programs_lock = PTHREAD_MUTEX_INITIALIZER;

static orc_program *orc_prog_of(orc_handle *h) {
    orc_program *r = NULL;
    pthread_mutex_lock(&g_programs_lock);
    for (int i = 0; i < ORC_MAX_PROGRAMS; i++)
        if (g_programs[i].h == h) r = g_programs[i].pr;
    pthread_mutex_unlock(&g_programs_lock);
    return r;
}

static int rec_is_array(int32_t kind) { return kind >= B200LP_REC_NODE_FLOW && kind <= B200LP_REC_STORAGE_PC; }

static void orc_free_program(orc_program *pr) {
    if (!pr) return;
    b200lp_program *p = &pr->p;
    free((void *)p->ts_days); free((void *)p->op_kind); free((void *)p->op_dst); free((void *)p->op_arg_ptr);
    free((void *)p->op_args); free((void *)p->table_rows); free((void *)p->table_cols); free((void *)p->table_axis);
    free((void *)p->table_offset); free((void *)p->table_data); free((void *)p->scen_index);
    free((void *)p->stflow_indptr); free((void *)p->stflow_col); free((void *)p->stflow_w);
    free((void *)p->initial_volume); free((void *)p->initial_pc);
    free((void *)p->rec_kind); free((void *)p->rec_row); free((void *)p->rec_src); free((void *)p->rec_ref); free((void *)p->rec_factor);
    free((void *)p->rec_indptr); free((void *)p->rec_col); free((void *)p->rec_w);
    if (pr->rec_out) for (int r = 0; r < p->n_recorders; r++) free(pr->rec_out[r]);
    free(pr->rec_out); free(pr->last_x);
    if (pr->roll_mem) for (int q = 0; q < pr->n_roll_storage; q++) free(pr->roll_mem[q]);
    free(pr->roll_mem); free(pr->roll_len); free(pr->roll_init);
    free(pr);
}

/* replace the contents of table `table` (same shape): b200lp_update_table */
int orc_update_table(void *hv, int32_t table, const double *data) {
    orc_handle *h = (orc_handle *)hv;
    orc_program *pr = orc_prog_of(h);
    if (!pr || !data || table < 0 || table >= pr->p.n_tables) return -1;
    const b200lp_program *p = &pr->p;
    memcpy((double *)p->table_data + p->table_offset[table], data,
           sizeof(double) * (size_t)p->table_rows[table] * p->table_cols[table]);
    return 0;
}

void orc_program_reset(void *hv) {
    orc_handle *h = (orc_handle *)hv;
    orc_program *pr = orc_prog_of(h);
    if (!pr) return;
    const b200lp_program *p = &pr->p;
    orc_reset(h, p->initial_volume);
    for (int32_t k = 0; k < h->S; k++) {
        h->sc[k].dead = 0;
        for (int32_t q = 0; q < h->s.n_storage; q++) h->sc[k].pc[q] = p->initial_pc[(size_t)q * h->S + k];
    }
    for (int r = 0; r < p->n_recorders; r++) {
        const size_t count = rec_is_array(p->rec_kind[r]) ? (size_t)p->n_steps * h->S : (size_t)h->S;
        for (size_t i = 0; i < count; i++) pr->rec_out[r][i] = (p->rec_kind[r] == B200LP_REC_MIN_VOLUME) ? INFINITY : 0.0;
    }
    for (int32_t q = 0; q < pr->n_roll_storage; q++)
        if (pr->roll_mem[q])
            for (size_t i = 0; i < (size_t)pr->roll_len[q] * h->S; i++) pr->roll_mem[q][i] = pr->roll_init[q];
    pr->n_failed = 0; pr->fail_scenario = -1; pr->fail_code = 0; pr->fail_step = -1;
}

int orc_load_program(void *hv, const b200lp_program *src) {
    orc_handle *h = (orc_handle *)hv;
    if (!h || !src) return -1;
    orc_program *pr = (orc_program *)calloc(1, sizeof(orc_program));
    pr->p = *src;
    b200lp_program *p = &pr->p;
    const int32_t S = h->S, nst = h->s.n_storage;
    p->ts_days = dup_mem(src->ts_days, sizeof(double) * src->n_steps);
    p->op_kind = dup_mem(src->op_kind, sizeof(int32_t) * src->n_ops);
    p->op_dst = dup_mem(src->op_dst, sizeof(int32_t) * src->n_ops);
    p->op_arg_ptr = dup_mem(src->op_arg_ptr, sizeof(int32_t) * (src->n_ops + 1));
    const int32_t nargs = src->n_ops ? src->op_arg_ptr[src->n_ops] : 0;
    p->op_args = dup_mem(src->op_args, sizeof(int32_t) * nargs);
    p->table_rows = dup_mem(src->table_rows, sizeof(int32_t) * src->n_tables);
    p->table_cols = dup_mem(src->table_cols, sizeof(int32_t) * src->n_tables);
    p->table_axis = dup_mem(src->table_axis, sizeof(int32_t) * src->n_tables);
    p->table_offset = dup_mem(src->table_offset, sizeof(int64_t) * src->n_tables);
    int64_t ndata = 0;
    for (int t = 0; t < src->n_tables; t++) {
        const int64_t end = src->table_offset[t] + (int64_t)src->table_rows[t] * src->table_cols[t];
        if (end > ndata) ndata = end;
    }
    p->table_data = dup_mem(src->table_data, sizeof(double) * (size_t)ndata);
    p->scen_index = dup_mem(src->scen_index, sizeof(int32_t) * (size_t)src->n_axes * S);
    p->stflow_indptr = dup_mem(src->stflow_indptr, sizeof(int32_t) * (nst + 1));
    const int32_t nsf = nst ? src->stflow_indptr[nst] : 0;
    p->stflow_col = dup_mem(src->stflow_col, sizeof(int32_t) * nsf);
    p->stflow_w = dup_mem(src->stflow_w, sizeof(double) * nsf);
    p->initial_volume = dup_mem(src->initial_volume, sizeof(double) * (size_t)nst * S);
    p->initial_pc = dup_mem(src->initial_pc, sizeof(double) * (size_t)nst * S);
    p->rec_kind = dup_mem(src->rec_kind, sizeof(int32_t) * src->n_recorders);
    p->rec_src = dup_mem(src->rec_src, sizeof(int32_t) * src->n_recorders);
    p->rec_row = dup_mem(src->rec_row, sizeof(int32_t) * src->n_recorders);
    p->rec_ref = dup_mem(src->rec_ref, sizeof(int32_t) * src->n_recorders);
    p->rec_factor = dup_mem(src->rec_factor, sizeof(double) * src->n_recorders);
    p->rec_indptr = dup_mem(src->rec_indptr, sizeof(int32_t) * (src->n_recorders + 1));
    const int32_t nrc = src->n_recorders ? src->rec_indptr[src->n_recorders] : 0;
    p->rec_col = dup_mem(src->rec_col, sizeof(int32_t) * nrc);
    p->rec_w = dup_mem(src->rec_w, sizeof(double) * nrc);
    pr->rec_out = (double **)calloc(src->n_recorders + 1, sizeof(double *));
    for (int r = 0; r < src->n_recorders; r++) {
        const size_t count = rec_is_array(src->rec_kind[r]) ? (size_t)src->n_steps * S : (size_t)S;
        pr->rec_out[r] = (double *)calloc(count, sizeof(double));
    }
    pr->last_x = (double *)calloc((size_t)h->n * S + 1, sizeof(double));
    pr->n_roll_storage = h->s.n_storage;
    pr->roll_mem = (double **)calloc(h->s.n_storage + 1, sizeof(double *));
    pr->roll_len = (int32_t *)calloc(h->s.n_storage + 1, sizeof(int32_t));
    pr->roll_init = (double *)calloc(h->s.n_storage + 1, sizeof(double));
    for (int32_t q = 0; q < h->s.n_storage; q++) {
        pr->roll_len[q] = src->st_roll_len ? src->st_roll_len[q] : 0;
        pr->roll_init[q] = (src->st_roll_init && pr->roll_len[q] > 0) ? src->st_roll_init[q] : 0.0;
        if (pr->roll_len[q] > 0) pr->roll_mem[q] = (double *)calloc((size_t)pr->roll_len[q] * S, sizeof(double));
    }
    p->st_roll_len = NULL; p->st_roll_init = NULL;  /* the copies above are the ones in use */
    pthread_mutex_init(&pr->lock, NULL);
    pr->loaded = 1;
    pthread_mutex_lock(&g_programs_lock);
    int placed = 0;
    for (int i = 0; i < ORC_MAX_PROGRAMS && !placed; i++) {
        if (g_programs[i].h == h) { orc_free_program(g_programs[i].pr); g_programs[i].pr = pr; placed = 1; }
    }
    for (int i = 0; i < ORC_MAX_PROGRAMS && !placed; i++) {
        if (g_programs[i].h == NULL) { g_programs[i].h = h; g_programs[i].pr = pr; placed = 1; }
    }
    pthread_mutex_unlock(&g_programs_lock);
    if (!placed) { orc_free_program(pr); return -2; }
    orc_program_reset(h);
    return 0;
}

void orc_unload_program(void *hv) {
    pthread_mutex_lock(&g_programs_lock);
    for (int i = 0; i < ORC_MAX_PROGRAMS; i++)
        if (g_programs[i].h == (orc_handle *)hv) { orc_free_program(g_programs[i].pr); g_programs[i].h = NULL; g_programs[i].pr = NULL; }
    pthread_mutex_unlock(&g_programs_lock);
}

/* Storage.get_current_pc (pywr/_core.pyx:1027-1039) */
static inline double clamp_pc(double pc) {
    if (pc > 1.0 || isnan(pc)) return 1.0;
    if (pc < 0.0) return 0.0;
    return pc;
}

static inline double pref(const orc_handle *h, int32_t ref, const double *p) {
    return ref >= 0 ? p[ref] : h->s.consts[-ref - 1];
}

static void orc_eval_ops(const orc_handle *h, const b200lp_program *pg, orc_scen *c, int32_t sidx, int32_t t,
                         double *p) {
    for (int32_t k = 0; k < pg->n_ops; k++) {
        const int32_t *a = pg->op_args + pg->op_arg_ptr[k];
        double v = 0.0;
        switch (pg->op_kind[k]) {
        case B200LP_OP_TABLE: {
            const int32_t tb = a[0];
            int64_t row = 0;
            if (pg->table_rows[tb] > 1) {
                row = (int64_t)t + a[1];
                if (row < 0) row = 0;
                if (row > pg->table_rows[tb] - 1) row = pg->table_rows[tb] - 1;
            }
            const int32_t ax = pg->table_axis[tb];
            const int32_t col = ax == B200LP_COL_NONE ? 0 : (ax == B200LP_COL_SCENARIO ? sidx : pg->scen_index[(size_t)ax * h->S + sidx]);
            v = pg->table_data[pg->table_offset[tb] + row * pg->table_cols[tb] + col];
        } break;
        case B200LP_OP_AGG: {
            const int32_t func = a[0], cnt = a[1];
            if (func == B200LP_AGG_PRODUCT) { v = 1.0; for (int i = 0; i < cnt; i++) v *= pref(h, a[2 + i], p); }
            else if (func == B200LP_AGG_SUM) { v = 0.0; for (int i = 0; i < cnt; i++) v += pref(h, a[2 + i], p); }
            else if (func == B200LP_AGG_MAX) { v = -INFINITY; for (int i = 0; i < cnt; i++) { double w = pref(h, a[2 + i], p); if (w > v) v = w; } }
            else if (func == B200LP_AGG_MIN) { v = INFINITY; for (int i = 0; i < cnt; i++) { double w = pref(h, a[2 + i], p); if (w < v) v = w; } }
            else { v = 0.0; for (int i = 0; i < cnt; i++) v += pref(h, a[2 + i], p); v /= cnt; }
        } break;
        case B200LP_OP_CONTROL_CURVE: {
            const int32_t cnt = a[1];
            const double pc = clamp_pc(c->pc[a[0]]);
            int32_t idx = cnt;
            for (int i = 0; i < cnt; i++) if (pc >= pref(h, a[2 + i], p)) { idx = i; break; }
            v = pref(h, a[2 + cnt + idx], p);
        } break;
        case B200LP_OP_CC_INDEX: {
            const int32_t cnt = a[1];
            const double pc = clamp_pc(c->pc[a[0]]);
            int32_t idx = cnt;
            for (int i = 0; i < cnt; i++) if (pc >= pref(h, a[2 + i], p)) { idx = i; break; }
            v = (double)idx;
        } break;
        case B200LP_OP_INDEXED: {
            int32_t idx = (int32_t)pref(h, a[0], p);
            if (idx < 0) idx = 0;
            if (idx > a[1] - 1) idx = a[1] - 1;
            v = pref(h, a[2 + idx], p);
        } break;
        case B200LP_OP_NEG: v = -pref(h, a[0], p); break;
        case B200LP_OP_MAX0: { const double x = pref(h, a[0], p), th = pref(h, a[1], p); v = x > th ? x : th; } break;
        case B200LP_OP_MIN0: { const double x = pref(h, a[0], p), th = pref(h, a[1], p); v = x < th ? x : th; } break;
        case B200LP_OP_DIV: v = pref(h, a[0], p) / pref(h, a[1], p); break;
        case B200LP_OP_STATE:  /* StorageParameter.value, _parameters.pyx:2032-2036 */
            v = a[0] == 1 ? c->vol[a[1]] : (a[0] == 3 ? clamp_pc(c->pc[a[1]]) : c->pc[a[1]]);
            break;
        case B200LP_OP_LAG:  /* FlowParameter / AbstractNode._prev_flow: the array recorder's sample of timestep t - 1 */
            /* pg is the first member of its orc_program, whose rec_out are the recorders' arrays */
            v = t < a[2] ? pref(h, a[1], p) : ((const orc_program *)pg)->rec_out[a[0]][(size_t)(t - a[2]) * h->S + sidx];
            break;
        case B200LP_OP_INTERP: {  /* AbstractInterpolatedParameter.value, pywr/parameters/parameters.py:156-158 (linear interp1d) */
            const int32_t cnt = a[2];
            const int32_t *xk = a + 3, *yk = a + 3 + cnt;
            const double x = a[0] == 1 ? c->vol[a[1]] : pref(h, a[1], p);
            const double x0 = pref(h, xk[0], p), xl = pref(h, xk[cnt - 1], p);
            if (x != x) v = NAN;
            else if (x < x0) v = pref(h, a[3 + 2 * cnt], p);
            else if (x > xl) v = pref(h, a[4 + 2 * cnt], p);
            else if (x == xl) v = pref(h, yk[cnt - 1], p);
            else {
                int j = 0;
                for (int i = 1; i < cnt - 1; i++) if (x >= pref(h, xk[i], p)) j = i;
                const double xa = pref(h, xk[j], p), ya = pref(h, yk[j], p);
                const double slope = (pref(h, yk[j + 1], p) - ya) / (pref(h, xk[j + 1], p) - xa);
                v = slope * (x - xa) + ya;
            }
        } break;
        case B200LP_OP_THRESHOLD: {  /* AbstractThresholdParameter.index / value, _thresholds.pyx:78-131 (no ratchet) */
            const double x = a[0] == 1 ? c->vol[a[1]] : pref(h, a[1], p), th = pref(h, a[3], p);
            int on;
            switch (a[2]) {
            case B200LP_PRED_LT: on = x < th; break;
            case B200LP_PRED_GT: on = x > th; break;
            case B200LP_PRED_LE: on = x <= th; break;
            case B200LP_PRED_GE: on = x >= th; break;
            default: on = x == th;
            }
            v = on ? pref(h, a[5], p) : pref(h, a[4], p);
        } break;
        case B200LP_OP_CC_INTERP: {  /* ControlCurveInterpolatedParameter.value, _control_curves.pyx:176-216 */
            const int32_t cnt = a[1];
            const int32_t *cv = a + 2, *vv = a + 2 + cnt;
            const double pc = clamp_pc(c->pc[a[0]]);
            double cc_prev = 1.0;
            int done = 0;
            for (int i = 0; i < cnt && !done; i++) {
                const double cc = pref(h, cv[i], p);
                if (pc >= cc) {
                    const double den = cc_prev - cc;
                    if (den == 0.0) v = pref(h, vv[i + 1], p);   /* ZeroDivisionError branch */
                    else { const double w = (pc - cc) / den; v = pref(h, vv[i], p) * w + pref(h, vv[i + 1], p) * (1.0 - w); }
                    done = 1;
                }
                cc_prev = cc;
            }
            if (!done) {
                const double den = cc_prev - 0.0;
                if (den == 0.0) v = pref(h, vv[cnt], p);
                else { const double w = (pc - 0.0) / den; v = pref(h, vv[cnt], p) * w + pref(h, vv[cnt + 1], p) * (1.0 - w); }
            }
        } break;
        case B200LP_OP_CC_PIECEWISE: {  /* ControlCurvePiecewiseInterpolatedParameter.value, :269-292 and _interpolate :10-24 */
            const int32_t cnt = a[1];
            const int32_t *cv = a + 2, *up = a + 2 + cnt, *lw = a + 2 + cnt + (cnt + 1);
            const double mn = pref(h, a[2 + cnt + 2 * (cnt + 1)], p), mx = pref(h, a[3 + cnt + 2 * (cnt + 1)], p);
            const double pc = clamp_pc(c->pc[a[0]]);
            double cc_prev = mx, lb = mn;
            int32_t band = cnt;
            for (int i = 0; i < cnt; i++) {
                const double cc = pref(h, cv[i], p);
                if (pc >= cc) { band = i; lb = cc; break; }
                cc_prev = cc;
            }
            const double ub = cc_prev, lv = pref(h, lw[band], p), uv = pref(h, up[band], p);
            if (pc < lb) v = lv;
            else if (pc > ub) v = uv;
            else if (ub == lb) v = lv;
            else { const double f = (pc - lb) / (ub - lb); v = lv + (uv - lv) * f; }
        } break;
        case B200LP_OP_STORAGE_RESET: {  /* node.before(): Storage._reset_storage_only, pywr/_core.pyx:1294-1333 */
            const int32_t q = a[0];
            v = pref(h, a[1], p);
            if (v == 1.0) {
                const double mxv = pref(h, h->s.st_maxv_ref[q], p);
                c->vol[q] = mxv;
                c->pc[q] = (mxv == 0.0) ? NAN : mxv / mxv;
            } else if (v == 2.0) {
                c->vol[q] = pref(h, a[2], p);
                c->pc[q] = pref(h, a[3], p);
            }
        } break;
        default: v = NAN;
        }
        p[pg->op_dst[k]] = v;
    }
}

typedef struct { orc_handle *h; orc_program *pr; int32_t s0, s1, t0, n; int64_t pivots, refactors; } orc_run_job;

static void *orc_run_worker(void *arg) {
    orc_run_job *J = (orc_run_job *)arg;
    orc_handle *h = J->h;
    orc_program *pr = J->pr;
    const b200lp_program *pg = &pr->p;
    const b200lp_structure *s = &h->s;
    const int32_t m = h->m, n = h->n, S = h->S, nst = s->n_storage;
    double *work = (double *)malloc(sizeof(double) * (2 * (size_t)(m + n) + 4 * n + 2 * m + (size_t)m * n + 8));
    double *x = work + 2 * (size_t)(m + n) + 3 * (size_t)n + m; /* ORC_WORK_X */
    double *ract = x + n;                                       /* row activities */
    int32_t *st_row = (int32_t *)calloc(nst + 1, sizeof(int32_t));
    for (int32_t i = 0; i < m; i++)
        if (s->row_kind[i] == B200LP_ROW_STORAGE) st_row[s->row_lb_ref[i]] = i;
    double *p = (double *)calloc(s->n_slots + 1, sizeof(double));
    J->pivots = 0; J->refactors = 0;
    for (int32_t k = J->s0; k < J->s1; k++) {
        orc_scen *c = &h->sc[k];
        for (int32_t t = J->t0; t < J->t0 + J->n && !c->dead; t++) {
            const double days = pg->ts_days[t];
            orc_eval_ops(h, pg, c, k, t, p);
            slot_view sv = {p, 1, 0};
            const int st = orc_solve_scenario(h, k, days, sv, NULL, NULL, NULL, work, &J->pivots, &J->refactors);
            if (st != B200LP_ST_OPTIMAL) {
                c->dead = 1;
                pthread_mutex_lock(&pr->lock);
                pr->n_failed++;
                if (pr->fail_scenario < 0 || k < pr->fail_scenario) { pr->fail_scenario = k; pr->fail_code = st; pr->fail_step = t; }
                pthread_mutex_unlock(&pr->lock);
                break;
            }
            /* after(): storage mass balance */
            for (int32_t q = 0; q < nst; q++) {
                if (s->st_kind[q] == B200LP_ST_KIND_VIRTUAL && !(pref(h, s->st_active_ref[q], p) != 0.0)) continue;
                double flow = 0.0;
                if (s->st_kind[q] == B200LP_ST_KIND_STORAGE) flow = ract[st_row[q]];
                else
                    for (int32_t e = pg->stflow_indptr[q]; e < pg->stflow_indptr[q + 1]; e++)
                        flow = fma(pg->stflow_w[e], x[pg->stflow_col[e]], flow);
                double vol = c->vol[q] + flow * days;
                const double mxv = pref(h, s->st_maxv_ref[q], p), mnv = pref(h, s->st_minv_ref[q], p);
                if (pr->roll_len[q] > 0) {  /* RollingVirtualStorage.after, pywr/_core.pyx:1522-1536 + :1345-1349 */
                    double *mem = pr->roll_mem[q] + (size_t)(t % pr->roll_len[q]) * S + k;
                    vol += *mem;
                    vol = fmin(mxv, fmax(mnv, vol));
                    *mem = -flow * days;
                }
                if (fabs(vol - mxv) < 1e-6) vol = mxv;
                if (fabs(vol - mnv) < 1e-6) vol = mnv;
                c->vol[q] = vol;
                c->pc[q] = (mxv == 0.0) ? NAN : vol / mxv;
            }
            /* recorders */
            for (int32_t r = 0; r < pg->n_recorders; r++) {
                const int32_t kind = pg->rec_kind[r];
                double *out = pr->rec_out[r];
                double value = 0.0;
                if (pg->rec_row[r] >= 0) value = ract[pg->rec_row[r]];
                else
                    for (int32_t e = pg->rec_indptr[r]; e < pg->rec_indptr[r + 1]; e++)
                        value = fma(pg->rec_w[e], x[pg->rec_col[e]], value);
                const double f = pg->rec_factor[r];
                switch (kind) {
                case B200LP_REC_NODE_FLOW: out[(size_t)t * S + k] = value * f; break;
                case B200LP_REC_NODE_DEFICIT: out[(size_t)t * S + k] = pref(h, pg->rec_ref[r], p) - value; break;
                case B200LP_REC_NODE_SUPPLIED: { const double mf = pref(h, pg->rec_ref[r], p); out[(size_t)t * S + k] = mf == 0.0 ? 1.0 : value / mf; } break;
                case B200LP_REC_NODE_CURTAIL: { const double mf = pref(h, pg->rec_ref[r], p); out[(size_t)t * S + k] = mf == 0.0 ? 0.0 : 1.0 - value / mf; } break;
                case B200LP_REC_STORAGE_VOLUME: out[(size_t)t * S + k] = c->vol[pg->rec_src[r]]; break;
                case B200LP_REC_STORAGE_PC: out[(size_t)t * S + k] = c->pc[pg->rec_src[r]]; break;
                case B200LP_REC_TOTAL_DEFICIT: out[k] += (pref(h, pg->rec_ref[r], p) - value) * days; break;
                case B200LP_REC_TOTAL_FLOW: out[k] += value * f * days; break;
                case B200LP_REC_MEAN_FLOW: out[k] += value * f; break;
                case B200LP_REC_DEFICIT_FREQ: if (fabs(value - pref(h, pg->rec_ref[r], p)) > 1e-6) out[k] += 1.0; break;
                case B200LP_REC_MIN_VOLUME: { const double v = c->vol[pg->rec_src[r]]; if (v < out[k]) out[k] = v; } break;
                default: break;
                }
            }
            if (t == J->t0 + J->n - 1)
                for (int32_t j = 0; j < n; j++) pr->last_x[(size_t)j * S + k] = x[j];
        }
    }
    free(work); free(p); free(st_row);
    return NULL;
}

int orc_run(void *hv, int32_t t0, int32_t nsteps) {
    orc_handle *h = (orc_handle *)hv;
    orc_program *pr = orc_prog_of(h);
    if (!pr || t0 < 0 || t0 + nsteps > pr->p.n_steps) return -1;
    int nt = h->threads < 1 ? 1 : h->threads;
    if (nt > h->S) nt = h->S;
    orc_run_job *jobs = (orc_run_job *)calloc(nt, sizeof(orc_run_job));
    pthread_t *tid = (pthread_t *)calloc(nt, sizeof(pthread_t));
    for (int t = 0; t < nt; t++) {
        jobs[t].h = h; jobs[t].pr = pr; jobs[t].t0 = t0; jobs[t].n = nsteps;
        jobs[t].s0 = (int32_t)((int64_t)h->S * t / nt);
        jobs[t].s1 = (int32_t)((int64_t)h->S * (t + 1) / nt);
    }
    for (int t = 1; t < nt; t++) pthread_create(&tid[t], NULL, orc_run_worker, &jobs[t]);
    orc_run_worker(&jobs[0]);
    for (int t = 1; t < nt; t++) pthread_join(tid[t], NULL);
    for (int t = 0; t < nt; t++) { h->pivots += jobs[t].pivots; h->refactors += jobs[t].refactors; }
    h->steps += nsteps;
    free(jobs); free(tid);
    return pr->n_failed;
}

void orc_program_status(void *hv, int32_t *n_failed, int32_t *scenario, int32_t *code, int32_t *timestep) {
    orc_program *pr = orc_prog_of((orc_handle *)hv);
    if (!pr) return;
    if (n_failed) *n_failed = pr->n_failed;
    if (scenario) *scenario = pr->fail_scenario;
    if (code) *code = pr->fail_code;
    if (timestep) *timestep = pr->fail_step;
}

int orc_fetch_recorder(void *hv, int32_t rec, double *out) {
    orc_handle *h = (orc_handle *)hv;
    orc_program *pr = orc_prog_of(h);
    if (!pr || rec < 0 || rec >= pr->p.n_recorders) return -1;
    const size_t count = rec_is_array(pr->p.rec_kind[rec]) ? (size_t)pr->p.n_steps * h->S : (size_t)h->S;
    memcpy(out, pr->rec_out[rec], sizeof(double) * count);
    return 0;
}

int orc_fetch_state(void *hv, double *col_flow, double *volume, double *pc) {
    orc_handle *h = (orc_handle *)hv;
    orc_program *pr = orc_prog_of(h);
    if (col_flow && pr) memcpy(col_flow, pr->last_x, sizeof(double) * (size_t)h->n * h->S);
    for (int32_t q = 0; q < h->s.n_storage; q++)
        for (int32_t k = 0; k < h->S; k++) {
            if (volume) volume[(size_t)q * h->S + k] = h->sc[k].vol[q];
            if (pc) pc[(size_t)q * h->S + k] = h->sc[k].pc[q];
        }
    return 0;
}
