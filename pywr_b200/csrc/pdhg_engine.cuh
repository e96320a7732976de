// pdhg_engine.cuh — batched restarted PDHG for networks whose dictionary does not fit the simplex
// kernels (BASELINE.json north_star: "for networks too large for shared memory there is a batched
// restarted-PDHG path, whose SpMM over scenario-contiguous vectors uses coalesced, vectorised 128-bit
// loads and shared-memory staging").
//
// What it replaces in the reference: CythonGLPKEdgeSolver._solve_scenario for every scenario of one
// timestep (pywr/solvers/cython_glpk.pyx:1649-1918; bound typing :66-95,:932-1019; costs :880-897;
// result extraction :1899-1911).  GLPK's simplex is replaced by a first-order method, so parity with
// the reference is by objective and feasibility residuals (tests/test_gpu_pdhg.py), not by vertex.
//
// Layout: every per-scenario vector is a stack of planes [item][ld] with the scenario index
// contiguous (ld = S rounded up to 2); one thread owns TWO neighbouring scenarios and moves them with
// one 128-bit load / store, a warp moves 512 contiguous bytes per plane row.  The constraint matrix is
// shared by all scenarios: each CTA stages the CSC (x update) or CSR (y update) slice of its chunk of
// columns / rows in shared memory once and every thread reads it as a broadcast.
//
// Algorithm (same steps, same names as oracle/pdhg_oracle.py):
//   1. structural presolve on the host (PdhgPlan): singleton rows -> column bounds, never-binding rows
//      dropped;  2. Ruiz + Pock-Chambolle scaling, ||A|| by power iteration;
//   3. reflected restarted Halpern PDHG, restart / convergence tests every `check` iterations on
//      deterministic two-stage reductions;  4. warm start from the scenario's previous timestep;
//   5. bound snapping and un-scaling, node flows.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/b200lp.h"

namespace b200lp {

constexpr double PD_BIG = 1e300;
constexpr int PD_THREADS = 128;   // 256 scenarios per CTA
constexpr int PD_CH = 8;          // columns / rows per CTA in the iteration kernels (dense phase)
constexpr int PD_CH_TAIL = 2;     // ... when few scenario pairs still iterate (latency-bound tail)
constexpr int PD_CHK = 64;        // columns / rows per CTA in the check kernels
constexpr int PD_MAXE = 512;      // matrix entries staged in shared memory per CTA
enum : int { Q_DX2 = 0, Q_DDX2, Q_POBJ, Q_XN2, Q_PR2, Q_DY2, Q_CROSS, Q_DDY2, Q_DOBJR, Q_DINFR, Q_YN2, Q_DRESC, Q_DOBJC, Q_COUNT };
// primal weight update at a restart: only when both iterates moved by more than this fraction of their norm since the
// last restart (a warm-started x that is already optimal barely moves: the ratio of two noise terms would blow omega up),
// and never further than this factor from the initial weight ||c|| / ||b||
constexpr double PD_OMEGA_MOVE = 1e-5, PD_OMEGA_SPAN = 1e4, PD_OMEGA_KICK = 4.0;

// omega after a restart (same rule as oracle/pdhg_oracle.py::solve_scaled): the PDLP movement rule when both iterates
// moved; a bounded kick when only one of them moved and the residual it has to repair dominates the error
__device__ __forceinline__ double pd_new_omega(const double omega, const double ddx, const double ddy, const double xn, const double yn,
                                               const double e_p, const double e_d, const double e_g, const double bnorm, const double cnorm) {
    const double om0 = (bnorm > 0.0 && cnorm > 0.0) ? cnorm / bnorm : 1.0;
    const bool mx = ddx > PD_OMEGA_MOVE * xn, my = ddy > PD_OMEGA_MOVE * yn;
    double om = omega;
    if (mx && my) om = exp(0.5 * log(ddy / ddx) + 0.5 * log(omega));
    else if (my && e_p >= fmax(e_d, e_g)) om = omega * PD_OMEGA_KICK;
    else if (mx && e_p < fmax(e_d, e_g)) om = omega / PD_OMEGA_KICK;
    return fmin(fmax(om, om0 / PD_OMEGA_SPAN), om0 * PD_OMEGA_SPAN);
}

struct PdhgDev {
    int m, n, m_r, N, n_slots, n_storage, cost_clip, S, ld, nch;
    const int *row_kind, *row_lb_ref, *row_ub_ref;
    const int *st_kind, *st_vol_ref, *st_minv_ref, *st_maxv_ref, *st_active_ref;
    const int *cost_indptr, *cost_ref;
    const double *cost_w;
    const int *flow_indptr, *flow_col;
    const double *flow_w;
    const double *consts;
    const int *keep;                    // [m_r] original row of each reduced row
    const int *sing_ptr, *sing_row;     // per column: singleton rows that bound it
    const double *sing_coef;
    const double *dr, *dc;              // row / column scaling
    const int *alias_root, *elim_row;   // [n] equality-row elimination: != j for a folded column; the row removed with it
    const int *al_ptr, *al_col;         // x_j = sum of al_ratio x_{al_col} (a surviving column: itself, ratio 1)
    const double *al_ratio;
    const int *cm_ptr, *cm_col;         // per column: itself and the columns folded into it (cost members)
    const double *cm_ratio;
    const int *csr_ptr, *csr_col;       // reduced, scaled matrix by rows
    const double *csr_val;
    const int *csc_ptr, *csc_row;       // ... and by columns
    const double *csc_val;
    const int *long_rows;               // rows with more than PD_LONG entries (aggregated nodes): summed by a whole warp in the tail kernel
    int n_long;
    double eta, tol;
    int max_iter, check;
};

struct PdhgState {
    double *x, *xa, *xb, *c, *lc, *uc, *tx;  // [n][ld]
    double *y, *ya, *lo, *hi, *ty;           // [m_r][ld]
    double *omega, *bnorm, *cnorm, *r0, *rprev, *objective;  // [ld]
    int *kbase, *done, *status, *valid, *action, *iters;     // [ld]
    double *vol;    // [n_storage][S] device-resident storage volumes (st_vol_ref == B200LP_REF_STATE)
    double *part;   // [Q_COUNT][nch][ld] partial sums of the check kernels
    int *it0;       // [1] iterations done so far in this solve
    int *n_active;  // [1] scenarios still iterating after the last check
    int *plist;     // [ld/2] scenario pairs with at least one scenario still iterating, ascending
    int *n_pairs;   // [1]
    unsigned long long *total_iters;  // [1] scenario-iterations since create
};

__device__ __forceinline__ double pd_ref(const PdhgDev &P, const double *__restrict__ slots, const int ref, const int s) {
    return ref >= 0 ? slots[(size_t)ref * P.S + s] : __ldg(&P.consts[-ref - 1]);
}
__device__ __forceinline__ double pd_clip8(const double v) { return fabs(v) < 1e-8 ? 0.0 : v; }
__device__ __forceinline__ double pd_fin(const double v) { return fabs(v) < PD_BIG ? v : 0.0; }

// (lb, ub) of ORIGINAL row i for scenario s: cython_glpk.pyx:1753-1840 (same typing as :932-1019)
__device__ __forceinline__ void pd_row_bounds(const PdhgDev &P, const double *__restrict__ slots, const double *__restrict__ vol,
                                              const int i, const int s, const double days, double &lo, double &hi) {
    if (__ldg(&P.row_kind[i]) == B200LP_ROW_FLOW) {
        lo = pd_clip8(pd_ref(P, slots, __ldg(&P.row_lb_ref[i]), s));
        hi = pd_clip8(pd_ref(P, slots, __ldg(&P.row_ub_ref[i]), s));
    } else {
        const int q = __ldg(&P.row_lb_ref[i]);
        const double maxv = pd_ref(P, slots, __ldg(&P.st_maxv_ref[q]), s), minv = pd_ref(P, slots, __ldg(&P.st_minv_ref[q]), s);
        const int vref = __ldg(&P.st_vol_ref[q]);
        const double v = vref == B200LP_REF_STATE ? vol[(size_t)q * P.S + s] : pd_ref(P, slots, vref, s);
        const bool active = __ldg(&P.st_kind[q]) == B200LP_ST_KIND_STORAGE || pd_ref(P, slots, __ldg(&P.st_active_ref[q]), s) != 0.0;
        if (maxv == minv) { lo = 0.0; hi = 0.0; return; }
        if (!active) { lo = -INFINITY; hi = INFINITY; return; }
        lo = pd_clip8(-fmax(v - minv, 0.0) / days);
        hi = pd_clip8(fmax(maxv - v, 0.0) / days);
        if (isnan(maxv) || isnan(minv) || isnan(v)) lo = NAN;
    }
    if (fabs(lo - hi) < 1e-8) hi = lo;
}

// ---- per-timestep assembly (one thread per scenario and row / column) --------------------------------
__global__ void pdhg_assemble_rows(const PdhgDev P, const PdhgState Z, const double *__restrict__ slots, const double days) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (s >= P.ld) return;
    double lo = 0.0, hi = 0.0;
    if (s < P.S) {
        pd_row_bounds(P, slots, Z.vol, __ldg(&P.keep[r]), s, days, lo, hi);
        if (isnan(lo) || isnan(hi)) atomicMax(&Z.status[s], B200LP_ST_NAN);
        else if (lo > hi) atomicMax(&Z.status[s], B200LP_ST_INFEASIBLE);
        const double d = __ldg(&P.dr[r]);
        lo *= d; hi *= d;
    }
    Z.lo[(size_t)r * P.ld + s] = lo;
    Z.hi[(size_t)r * P.ld + s] = hi;
}

__global__ void pdhg_assemble_cols(const PdhgDev P, const PdhgState Z, const double *__restrict__ slots, const double days) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (s >= P.ld) return;
    double lc = 0.0, uc = 0.0, c = 0.0;
    if (s < P.S) {
        uc = INFINITY;
        bool bad = false;
        for (int e = __ldg(&P.sing_ptr[j]); e < __ldg(&P.sing_ptr[j + 1]); e++) {
            double l, u;
            pd_row_bounds(P, slots, Z.vol, __ldg(&P.sing_row[e]), s, days, l, u);
            bad |= isnan(l) || isnan(u);
            const double a = __ldg(&P.sing_coef[e]);
            const double l2 = a > 0.0 ? l / a : u / a, u2 = a > 0.0 ? u / a : l / a;
            lc = fmax(lc, l2); uc = fmin(uc, u2);
        }
        for (int g = __ldg(&P.cm_ptr[j]); g < __ldg(&P.cm_ptr[j + 1]); g++) {   // the column and every column folded into it
            const int col = __ldg(&P.cm_col[g]);
            double acc = 0.0;
            for (int e = __ldg(&P.cost_indptr[col]); e < __ldg(&P.cost_indptr[col + 1]); e++)
                acc += __ldg(&P.cost_w[e]) * pd_ref(P, slots, __ldg(&P.cost_ref[e]), s);
            if (P.cost_clip && fabs(acc) < 1e-8) acc = 0.0;
            c += __ldg(&P.cm_ratio[g]) * acc;
        }
        if (__ldg(&P.alias_root[j]) != j) uc = 0.0;   // folded column: fixed at 0 here, rebuilt in pdhg_unscale
        if (bad || isnan(c)) atomicMax(&Z.status[s], B200LP_ST_NAN);
        else if (lc > uc + 1e-9) atomicMax(&Z.status[s], B200LP_ST_INFEASIBLE);
        uc = fmax(uc, lc);
        const double d = __ldg(&P.dc[j]);
        c *= d; lc /= d; uc /= d;
    }
    Z.c[(size_t)j * P.ld + s] = c;
    Z.lc[(size_t)j * P.ld + s] = lc;
    Z.uc[(size_t)j * P.ld + s] = uc;
}

// norms, cold / warm start, anchors (one thread per scenario; plane rows are read coalesced)
__global__ void pdhg_prepare(const PdhgDev P, const PdhgState Z) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.ld) return;
    const size_t ld = P.ld;
    if (s == 0) { *Z.it0 = 0; *Z.n_active = 0; }
    if (s >= P.S || Z.status[s] != B200LP_ST_OPTIMAL) {
        Z.done[s] = 1; Z.iters[s] = 0;
        return;
    }
    double b2 = 0.0, c2 = 0.0;
    for (int r = 0; r < P.m_r; r++) {
        const double b = fmax(fabs(pd_fin(Z.lo[r * ld + s])), fabs(pd_fin(Z.hi[r * ld + s])));
        b2 = fma(b, b, b2);
    }
    const bool warm = Z.valid[s] != 0;
    for (int j = 0; j < P.n; j++) {
        const double c = Z.c[j * ld + s], l = Z.lc[j * ld + s], u = Z.uc[j * ld + s];
        c2 = fma(c, c, c2);
        b2 = fma(pd_fin(l), pd_fin(l), b2);
        b2 = fma(pd_fin(u), pd_fin(u), b2);
        const double x = warm ? Z.x[j * ld + s] : 0.0;
        const double xc = fmin(fmax(x, l), u);
        Z.x[j * ld + s] = xc; Z.xa[j * ld + s] = xc;
    }
    for (int r = 0; r < P.m_r; r++) {
        const double y = warm ? Z.y[r * ld + s] : 0.0;
        Z.y[r * ld + s] = y; Z.ya[r * ld + s] = y;
    }
    const double bnorm = sqrt(b2), cnorm = sqrt(c2);
    Z.bnorm[s] = bnorm; Z.cnorm[s] = cnorm;
    if (!warm) Z.omega[s] = (bnorm > 0.0 && cnorm > 0.0) ? cnorm / bnorm : 1.0;
    Z.valid[s] = 1;
    Z.kbase[s] = 0; Z.r0[s] = -1.0; Z.rprev[s] = INFINITY;
    Z.done[s] = 0; Z.iters[s] = 0; Z.action[s] = 0;
}

// shared-memory staging of the matrix slice [e0, e1) of a chunk (entries beyond PD_MAXE stay in L2)
struct PdStage {
    int ptr[PD_CHK + 1];
    int idx[PD_MAXE];
    double val[PD_MAXE];
};
template <int CH>
__device__ __forceinline__ void pd_stage(PdStage &sh, const int *__restrict__ ptr, const int *__restrict__ idx,
                                         const double *__restrict__ val, const int i0, const int cnt) {
    for (int k = threadIdx.x; k <= cnt; k += blockDim.x) sh.ptr[k] = __ldg(&ptr[i0 + k]);
    __syncthreads();
    const int e0 = sh.ptr[0], ne = min(sh.ptr[cnt] - e0, PD_MAXE);
    for (int k = threadIdx.x; k < ne; k += blockDim.x) { sh.idx[k] = __ldg(&idx[e0 + k]); sh.val[k] = __ldg(&val[e0 + k]); }
    __syncthreads();
}
#define PD_ENTRY(sh, gidx, gval, e, e0, I, A)                                      \
    do {                                                                          \
        const int _k = (e) - (e0);                                                \
        if (_k < PD_MAXE) { I = sh.idx[_k]; A = sh.val[_k]; }                     \
        else { I = __ldg(&(gidx)[e]); A = __ldg(&(gval)[e]); }                    \
    } while (0)

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, const double2 v) { *reinterpret_cast<double2 *>(p) = v; }
__device__ __forceinline__ double pd_clamp(const double v, const double lo, const double hi) { return fmin(fmax(v, lo), hi); }

// ---- x half-step:  xn = clip(x - tau (c - A^T y));  xb = 2 xn - x;  x <- w xb + (1 - w) xa -------------
template <int CH>
__global__ void __launch_bounds__(PD_THREADS) pdhg_x_kernel(const PdhgDev P, const PdhgState Z, const int j_in_batch) {
    __shared__ PdStage sh;
    const int np = *Z.n_pairs;
    if ((int)(blockIdx.x * blockDim.x) >= np) return;  // converged scenarios cost nothing: pairs are compacted
    const int j0 = blockIdx.y * CH, cnt = min(CH, P.n - j0);
    pd_stage<CH>(sh, P.csc_ptr, P.csc_row, P.csc_val, j0, cnt);
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= np) return;
    const int s0 = 2 * Z.plist[pt];
    const int2 dn = *reinterpret_cast<const int2 *>(Z.done + s0);
    const int2 kb = *reinterpret_cast<const int2 *>(Z.kbase + s0);
    const int it = *Z.it0 + j_in_batch;
    const double2 om = ld2(Z.omega + s0);
    const double tau0 = P.eta / om.x, tau1 = P.eta / om.y;
    const double k0 = (double)(it - kb.x), k1 = (double)(it - kb.y);
    const double w0 = (k0 + 1.0) / (k0 + 2.0), w1 = (k1 + 1.0) / (k1 + 2.0);
    const size_t ld = P.ld;
    const int e0 = sh.ptr[0];
#pragma unroll 4
    for (int jj = 0; jj < cnt; jj++) {
        const size_t at = (size_t)(j0 + jj) * ld + s0;
        double2 g = ld2(Z.c + at);
        const double2 x = ld2(Z.x + at), l = ld2(Z.lc + at), u = ld2(Z.uc + at), a = ld2(Z.xa + at);
        for (int e = sh.ptr[jj]; e < sh.ptr[jj + 1]; e++) {
            int r; double v;
            PD_ENTRY(sh, P.csc_row, P.csc_val, e, e0, r, v);
            const double2 y = ld2(Z.y + (size_t)r * ld + s0);
            g.x = fma(-v, y.x, g.x); g.y = fma(-v, y.y, g.y);
        }
        const double xn0 = pd_clamp(fma(-tau0, g.x, x.x), l.x, u.x), xn1 = pd_clamp(fma(-tau1, g.y, x.y), l.y, u.y);
        double2 xb, xo;
        xb.x = 2.0 * xn0 - x.x; xb.y = 2.0 * xn1 - x.y;
        xo.x = dn.x ? x.x : fma(w0, xb.x, (1.0 - w0) * a.x);
        xo.y = dn.y ? x.y : fma(w1, xb.y, (1.0 - w1) * a.y);
        st2(Z.xb + at, xb);
        st2(Z.x + at, xo);
    }
}

// ---- y half-step:  v = y - sigma A xb;  yn = v - clip(v, -sigma hi, -sigma lo);  y <- w (2 yn - y) + (1 - w) ya
template <int CH>
__global__ void __launch_bounds__(PD_THREADS) pdhg_y_kernel(const PdhgDev P, const PdhgState Z, const int j_in_batch) {
    __shared__ PdStage sh;
    const int np = *Z.n_pairs;
    if ((int)(blockIdx.x * blockDim.x) >= np) return;
    const int i0 = blockIdx.y * CH, cnt = min(CH, P.m_r - i0);
    pd_stage<CH>(sh, P.csr_ptr, P.csr_col, P.csr_val, i0, cnt);
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= np) return;
    const int s0 = 2 * Z.plist[pt];
    const int2 dn = *reinterpret_cast<const int2 *>(Z.done + s0);
    const int2 kb = *reinterpret_cast<const int2 *>(Z.kbase + s0);
    const int it = *Z.it0 + j_in_batch;
    const double2 om = ld2(Z.omega + s0);
    const double sg0 = P.eta * om.x, sg1 = P.eta * om.y;
    const double k0 = (double)(it - kb.x), k1 = (double)(it - kb.y);
    const double w0 = (k0 + 1.0) / (k0 + 2.0), w1 = (k1 + 1.0) / (k1 + 2.0);
    const size_t ld = P.ld;
    const int e0 = sh.ptr[0];
#pragma unroll 4
    for (int ii = 0; ii < cnt; ii++) {
        const size_t at = (size_t)(i0 + ii) * ld + s0;
        const double2 y = ld2(Z.y + at), lo = ld2(Z.lo + at), hi = ld2(Z.hi + at), a = ld2(Z.ya + at);
        double2 ax; ax.x = 0.0; ax.y = 0.0;
        for (int e = sh.ptr[ii]; e < sh.ptr[ii + 1]; e++) {
            int c; double v;
            PD_ENTRY(sh, P.csr_col, P.csr_val, e, e0, c, v);
            const double2 xb = ld2(Z.xb + (size_t)c * ld + s0);
            ax.x = fma(v, xb.x, ax.x); ax.y = fma(v, xb.y, ax.y);
        }
        const double v0 = fma(-sg0, ax.x, y.x), v1 = fma(-sg1, ax.y, y.y);
        const double yn0 = v0 - pd_clamp(v0, -sg0 * hi.x, -sg0 * lo.x), yn1 = v1 - pd_clamp(v1, -sg1 * hi.y, -sg1 * lo.y);
        double2 yo;
        yo.x = dn.x ? y.x : fma(w0, 2.0 * yn0 - y.x, (1.0 - w0) * a.x);
        yo.y = dn.y ? y.y : fma(w1, 2.0 * yn1 - y.y, (1.0 - w1) * a.y);
        st2(Z.y + at, yo);
    }
}

// ---- check, pass 1 (columns): tx = xn = T_x(z); partial sums |xn - x|^2, |xn - xa|^2, c.xn ------------
__global__ void __launch_bounds__(PD_THREADS) pdhg_chk_cols1(const PdhgDev P, const PdhgState Z) {
    __shared__ PdStage sh;
    const int np = *Z.n_pairs;
    if ((int)(blockIdx.x * blockDim.x) >= np) return;
    const int j0 = blockIdx.y * PD_CHK, cnt = min(PD_CHK, P.n - j0);
    pd_stage<PD_CHK>(sh, P.csc_ptr, P.csc_row, P.csc_val, j0, cnt);
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= np) return;
    const int s0 = 2 * Z.plist[pt];
    const size_t ld = P.ld;
    const double2 om = ld2(Z.omega + s0);
    const double tau0 = P.eta / om.x, tau1 = P.eta / om.y;
    double2 dx2 = {0.0, 0.0}, ddx2 = {0.0, 0.0}, pobj = {0.0, 0.0}, xn2 = {0.0, 0.0};
    const int e0 = sh.ptr[0];
    for (int jj = 0; jj < cnt; jj++) {
        const size_t at = (size_t)(j0 + jj) * ld + s0;
        const double2 c = ld2(Z.c + at);
        double2 g = c;
        const double2 x = ld2(Z.x + at), l = ld2(Z.lc + at), u = ld2(Z.uc + at), a = ld2(Z.xa + at);
        for (int e = sh.ptr[jj]; e < sh.ptr[jj + 1]; e++) {
            int r; double v;
            PD_ENTRY(sh, P.csc_row, P.csc_val, e, e0, r, v);
            const double2 y = ld2(Z.y + (size_t)r * ld + s0);
            g.x = fma(-v, y.x, g.x); g.y = fma(-v, y.y, g.y);
        }
        double2 xn;
        xn.x = pd_clamp(fma(-tau0, g.x, x.x), l.x, u.x); xn.y = pd_clamp(fma(-tau1, g.y, x.y), l.y, u.y);
        st2(Z.tx + at, xn);
        dx2.x = fma(xn.x - x.x, xn.x - x.x, dx2.x); dx2.y = fma(xn.y - x.y, xn.y - x.y, dx2.y);
        ddx2.x = fma(xn.x - a.x, xn.x - a.x, ddx2.x); ddx2.y = fma(xn.y - a.y, xn.y - a.y, ddx2.y);
        pobj.x = fma(c.x, xn.x, pobj.x); pobj.y = fma(c.y, xn.y, pobj.y);
        xn2.x = fma(xn.x, xn.x, xn2.x); xn2.y = fma(xn.y, xn.y, xn2.y);
    }
    const size_t pstride = (size_t)P.nch * ld, pat = (size_t)blockIdx.y * ld + s0;
    st2(Z.part + Q_DX2 * pstride + pat, dx2);
    st2(Z.part + Q_DDX2 * pstride + pat, ddx2);
    st2(Z.part + Q_POBJ * pstride + pat, pobj);
    st2(Z.part + Q_XN2 * pstride + pat, xn2);
}

// ---- check, pass 2 (rows): ty = yn = T_y(z); primal residual, fixed-point terms, dual objective (rows) ---
__global__ void __launch_bounds__(PD_THREADS) pdhg_chk_rows(const PdhgDev P, const PdhgState Z) {
    __shared__ PdStage sh;
    const int np = *Z.n_pairs;
    if ((int)(blockIdx.x * blockDim.x) >= np) return;
    const int i0 = blockIdx.y * PD_CHK, cnt = min(PD_CHK, P.m_r - i0);
    pd_stage<PD_CHK>(sh, P.csr_ptr, P.csr_col, P.csr_val, i0, cnt);
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= np) return;
    const int s0 = 2 * Z.plist[pt];
    const size_t ld = P.ld;
    const double2 om = ld2(Z.omega + s0);
    const double sg[2] = {P.eta * om.x, P.eta * om.y};
    double acc[7][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};  // pr2 dy2 cross ddy2 dobj dinf2 yn2
    const int e0 = sh.ptr[0];
    for (int ii = 0; ii < cnt; ii++) {
        const size_t at = (size_t)(i0 + ii) * ld + s0;
        const double2 y2 = ld2(Z.y + at), lo2 = ld2(Z.lo + at), hi2 = ld2(Z.hi + at), a2 = ld2(Z.ya + at);
        double axn[2] = {0.0, 0.0}, ax[2] = {0.0, 0.0};
        for (int e = sh.ptr[ii]; e < sh.ptr[ii + 1]; e++) {
            int c; double v;
            PD_ENTRY(sh, P.csr_col, P.csr_val, e, e0, c, v);
            const double2 xn = ld2(Z.tx + (size_t)c * ld + s0), x = ld2(Z.x + (size_t)c * ld + s0);
            axn[0] = fma(v, xn.x, axn[0]); axn[1] = fma(v, xn.y, axn[1]);
            ax[0] = fma(v, x.x, ax[0]); ax[1] = fma(v, x.y, ax[1]);
        }
        const double y[2] = {y2.x, y2.y}, lo[2] = {lo2.x, lo2.y}, hi[2] = {hi2.x, hi2.y}, ya[2] = {a2.x, a2.y};
        double yn[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double axb = 2.0 * axn[q] - ax[q];
            const double v = fma(-sg[q], axb, y[q]);
            yn[q] = v - pd_clamp(v, -sg[q] * hi[q], -sg[q] * lo[q]);
            const double viol = axn[q] - pd_clamp(axn[q], lo[q], hi[q]);
            const double dy = yn[q] - y[q];
            const double yp = fmax(yn[q], 0.0), ym = fmax(-yn[q], 0.0);
            acc[0][q] = fma(viol, viol, acc[0][q]);
            acc[1][q] = fma(dy, dy, acc[1][q]);
            acc[2][q] = fma(dy, axn[q] - ax[q], acc[2][q]);
            acc[3][q] = fma(yn[q] - ya[q], yn[q] - ya[q], acc[3][q]);
            acc[4][q] += pd_fin(lo[q]) * yp - pd_fin(hi[q]) * ym;
            const double il = fabs(lo[q]) < PD_BIG ? 0.0 : yp, ih = fabs(hi[q]) < PD_BIG ? 0.0 : ym;
            acc[5][q] = fma(il, il, fma(ih, ih, acc[5][q]));
            acc[6][q] = fma(yn[q], yn[q], acc[6][q]);
        }
        double2 o; o.x = yn[0]; o.y = yn[1];
        st2(Z.ty + at, o);
    }
    const size_t pstride = (size_t)P.nch * ld, pat = (size_t)blockIdx.y * ld + s0;
    const int qs[7] = {Q_PR2, Q_DY2, Q_CROSS, Q_DDY2, Q_DOBJR, Q_DINFR, Q_YN2};
#pragma unroll
    for (int k = 0; k < 7; k++) { double2 o; o.x = acc[k][0]; o.y = acc[k][1]; st2(Z.part + qs[k] * pstride + pat, o); }
}

// ---- check, pass 3 (columns): reduced costs of yn: dual residual and dual objective (columns) -----------
__global__ void __launch_bounds__(PD_THREADS) pdhg_chk_cols2(const PdhgDev P, const PdhgState Z) {
    __shared__ PdStage sh;
    const int np = *Z.n_pairs;
    if ((int)(blockIdx.x * blockDim.x) >= np) return;
    const int j0 = blockIdx.y * PD_CHK, cnt = min(PD_CHK, P.n - j0);
    pd_stage<PD_CHK>(sh, P.csc_ptr, P.csc_row, P.csc_val, j0, cnt);
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= np) return;
    const int s0 = 2 * Z.plist[pt];
    const size_t ld = P.ld;
    double dres[2] = {0.0, 0.0}, dobj[2] = {0.0, 0.0};
    const int e0 = sh.ptr[0];
    for (int jj = 0; jj < cnt; jj++) {
        const size_t at = (size_t)(j0 + jj) * ld + s0;
        double2 g = ld2(Z.c + at);
        const double2 l2 = ld2(Z.lc + at), u2 = ld2(Z.uc + at);
        for (int e = sh.ptr[jj]; e < sh.ptr[jj + 1]; e++) {
            int r; double v;
            PD_ENTRY(sh, P.csc_row, P.csc_val, e, e0, r, v);
            const double2 y = ld2(Z.ty + (size_t)r * ld + s0);
            g.x = fma(-v, y.x, g.x); g.y = fma(-v, y.y, g.y);
        }
        const double rc[2] = {g.x, g.y}, l[2] = {l2.x, l2.y}, u[2] = {u2.x, u2.y};
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double rp = fmax(rc[q], 0.0), rm = fmax(-rc[q], 0.0);
            const double il = fabs(l[q]) < PD_BIG ? 0.0 : rp, iu = fabs(u[q]) < PD_BIG ? 0.0 : rm;
            dres[q] = fma(il, il, fma(iu, iu, dres[q]));
            dobj[q] += pd_fin(l[q]) * rp - pd_fin(u[q]) * rm;
        }
    }
    const size_t pstride = (size_t)P.nch * ld, pat = (size_t)blockIdx.y * ld + s0;
    double2 o; o.x = dres[0]; o.y = dres[1]; st2(Z.part + Q_DRESC * pstride + pat, o);
    o.x = dobj[0]; o.y = dobj[1]; st2(Z.part + Q_DOBJC * pstride + pat, o);
}

__global__ void pdhg_advance(const PdhgState Z, const int check) { *Z.it0 += check; *Z.n_active = 0; }

// ---- decision per scenario: converged / restart / continue (fixed summation order: deterministic) -------
__global__ void pdhg_decide(const PdhgDev P, const PdhgState Z, const int nch_cols, const int nch_rows) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.ld) return;
    if (Z.done[s]) { Z.action[s] = 0; return; }
    const size_t ld = P.ld, pstride = (size_t)P.nch * ld;
    double q[Q_COUNT];
#pragma unroll
    for (int k = 0; k < Q_COUNT; k++) {
        const bool rows = (k >= Q_PR2 && k <= Q_YN2);
        const int nc = rows ? nch_rows : nch_cols;
        double acc = 0.0;
        for (int ch = 0; ch < nc; ch++) acc += Z.part[k * pstride + (size_t)ch * ld + s];
        q[k] = acc;
    }
    const double omega = Z.omega[s], tau = P.eta / omega, sig = P.eta * omega;
    const int it = *Z.it0;
    const double r = sqrt(fmax(q[Q_DX2] / tau - 2.0 * q[Q_CROSS] + q[Q_DY2] / sig, 0.0));
    const double pobj = q[Q_POBJ], dobj = q[Q_DOBJR] + q[Q_DOBJC];
    const double e_p = sqrt(q[Q_PR2]) / (1.0 + Z.bnorm[s]), e_d = sqrt(q[Q_DINFR] + q[Q_DRESC]) / (1.0 + Z.cnorm[s]);
    const double e_g = fabs(pobj - dobj) / (1.0 + fabs(pobj) + fabs(dobj));
    const double rel = fmax(fmax(e_p, e_d), e_g);
    double r0 = Z.r0[s];
    if (r0 < 0.0) r0 = r;
    const int k = it - Z.kbase[s];
    int action = 0;
    if (rel < P.tol || it >= P.max_iter || !(rel == rel)) {
        action = 2;
        Z.done[s] = 1;
        Z.status[s] = rel < P.tol ? B200LP_ST_OPTIMAL : (rel == rel ? B200LP_ST_ITERLIMIT : B200LP_ST_NAN);
        Z.iters[s] = it;
        atomicAdd(Z.total_iters, (unsigned long long)it);
    } else {
        if (r <= 0.2 * r0 || (r <= 0.8 * r0 && r > Z.rprev[s]) || (double)k >= 0.36 * (double)it) {
            action = 1;
            Z.omega[s] = pd_new_omega(omega, sqrt(q[Q_DDX2]), sqrt(q[Q_DDY2]), sqrt(q[Q_XN2]), sqrt(q[Q_YN2]), e_p, e_d, e_g,
                                      Z.bnorm[s], Z.cnorm[s]);
            Z.kbase[s] = it;
            r0 = -1.0;
        }
        atomicAdd(Z.n_active, 1);
    }
    Z.r0[s] = r0; Z.rprev[s] = r;
    Z.action[s] = action;
}

// restart / convergence: z := anchor := T(z)
__global__ void __launch_bounds__(PD_THREADS) pdhg_apply(const PdhgDev P, const PdhgState Z) {
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= *Z.n_pairs) return;
    const int s0 = 2 * Z.plist[pt];
    const int2 ac = *reinterpret_cast<const int2 *>(Z.action + s0);
    if (!ac.x && !ac.y) return;
    const size_t ld = P.ld;
    const int i0 = blockIdx.y * PD_CHK;
    for (int ii = 0; ii < PD_CHK; ii++) {
        const int i = i0 + ii;
        if (i >= P.n + P.m_r) return;
        const bool col = i < P.n;
        const size_t at = (size_t)(col ? i : i - P.n) * ld + s0;
        double *z = col ? Z.x : Z.y, *za = col ? Z.xa : Z.ya;
        const double *t = col ? Z.tx : Z.ty;
        const double2 tn = ld2(t + at);
        double2 zo = ld2(z + at), ao = ld2(za + at);
        if (ac.x) { zo.x = tn.x; ao.x = tn.x; }
        if (ac.y) { zo.y = tn.y; ao.y = tn.y; }
        st2(z + at, zo); st2(za + at, ao);
    }
}

// ascending list of the scenario pairs that still iterate (one CTA; block-wide exclusive scan)
__global__ void __launch_bounds__(1024) pdhg_compact(const PdhgDev P, const PdhgState Z) {
    __shared__ int warp_sum[32];
    __shared__ int base;
    const int npairs = P.ld / 2;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int p0 = 0; p0 < npairs; p0 += 1024) {
        const int p = p0 + threadIdx.x;
        const int flag = (p < npairs && !(Z.done[2 * p] && Z.done[2 * p + 1])) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_sum[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int k = 0; k < w; k++) off += warp_sum[k];
        if (flag) Z.plist[off + __popc(bal & ((1u << lane) - 1u))] = p;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 32; k++) t += warp_sum[k]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *Z.n_pairs = base;
}

// ------------------------------------------------------------------------------------------------------
// Tail: ONE CTA PER SCENARIO, every vector of the scenario in shared memory, the whole remaining solve
// (iterations, convergence tests, restarts) inside one launch.  Used when only a few scenarios still
// iterate: on the streaming path they cost a full launch pair per iteration (latency bound, scattered
// 16-byte accesses); here an iteration is two passes over shared memory with the shared matrix read
// through L1/L2.  Same arithmetic per element as the streaming kernels; the reductions of the check are
// block-wide (different summation order, same quantities).
// ------------------------------------------------------------------------------------------------------
constexpr int PD_TAIL_THREADS = 512;
constexpr int PD_TAIL_CPT = 6, PD_TAIL_RPT = 3;  // columns / rows owned by a thread: n <= 3072, m_r <= 1536
constexpr int PD_TAIL_CE = 3;                    // matrix entries per owned column cached in registers (rows: CSR in shared memory)
constexpr int PD_LONG = 12, PD_MAX_LONG = 64;     // "long" rows and how many of them the tail kernel supports    // matrix entries per owned column / row cached in registers
__host__ __device__ inline size_t pd_tail_smem_bytes(int n, int m_r, int nnz) {
    // vectors, reduction scratch, long-row sums, then the CSR of the shared matrix (values + 16-bit column indices)
    return sizeof(double) * ((size_t)6 * n + (size_t)5 * m_r + 32 * Q_COUNT + Q_COUNT + 8 + 2 * PD_MAX_LONG + (size_t)nnz) +
           sizeof(unsigned short) * (size_t)(nnz + 8);
}
inline bool pd_tail_fits(int n, int m_r, int nnz, int n_long) {
    return n_long <= PD_MAX_LONG && n <= 65535 && n <= PD_TAIL_CPT * PD_TAIL_THREADS && m_r <= PD_TAIL_RPT * PD_TAIL_THREADS &&
           pd_tail_smem_bytes(n, m_r, nnz) <= 227 * 1024;
}

// the owned slices of the shared matrix, cached in registers for the whole launch (entries beyond the
// cached ones - e.g. the 200-entry licence row - are read through L1/L2)
struct PdTailCols {
    int row[PD_TAIL_CPT][PD_TAIL_CE];
    double val[PD_TAIL_CPT][PD_TAIL_CE];
    unsigned more;  // bit k: column slot k has entries beyond the cached ones (rare)
};
struct PdTailRows {
    int beg[PD_TAIL_RPT], end[PD_TAIL_RPT];  // entry range in the shared-memory CSR (empty for long rows)
    int lidx[PD_TAIL_RPT];                   // >= 0: the row is a long one, its sums come from the cooperative pass
};

// sums of the long rows, one warp per row: out[2r] = A_i . u, out[2r+1] = A_i . v (lanes stride over the entries)
__device__ __forceinline__ void pd_tail_long_rows(const PdhgDev &P, const double *rval, const unsigned short *rcol, const double *u,
                                                  const double *v, double *out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < P.n_long; r += PD_TAIL_THREADS / 32) {
        const int i = __ldg(&P.long_rows[r]);
        double a = 0.0, b = 0.0;
        for (int e = __ldg(&P.csr_ptr[i]) + lane; e < __ldg(&P.csr_ptr[i + 1]); e += 32) {
            const int cc = rcol[e];
            const double w = rval[e];
            a = fma(w, u[cc], a); b = fma(w, v[cc], b);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
        if (lane == 0) { out[2 * r] = a; out[2 * r + 1] = b; }
    }
}

// g = g0 - sum_e val_e * v[row_e] over column slot k = column j (entries in matrix order, as in the streaming kernels)
__device__ __forceinline__ double pd_tail_coldot(const PdhgDev &P, const PdTailCols &C, const int k, const int j, const double g0,
                                                 const double *v) {
    double g = g0;
#pragma unroll
    for (int e = 0; e < PD_TAIL_CE; e++) g = fma(-C.val[k][e], v[C.row[k][e]], g);   // absent entries: value 0, row 0
    if (C.more & (1u << k))
        for (int e = __ldg(&P.csc_ptr[j]) + PD_TAIL_CE; e < __ldg(&P.csc_ptr[j + 1]); e++) g = fma(-__ldg(&P.csc_val[e]), v[__ldg(&P.csc_row[e])], g);
    return g;
}

__global__ void __launch_bounds__(PD_TAIL_THREADS, 1) pdhg_tail_kernel(const PdhgDev P, const PdhgState Z) {
    extern __shared__ __align__(16) unsigned char pd_tail_raw[];
    const int b = blockIdx.x;
    if ((b >> 1) >= *Z.n_pairs) return;
    const int s = 2 * Z.plist[b >> 1] + (b & 1);
    if (Z.done[s]) return;
    const int n = P.n, m = P.m_r, tid = threadIdx.x;
    constexpr int NT = PD_TAIL_THREADS;
    double *x = reinterpret_cast<double *>(pd_tail_raw), *xa = x + n, *xb = xa + n, *c = xb + n, *lc = c + n, *uc = lc + n;
    double *y = uc + n, *ya = y + m, *lo = ya + m, *hi = lo + m, *ty = hi + m;
    double *red = ty + m;              // [32][Q_COUNT] per-warp partial sums
    double *tot = red + 32 * Q_COUNT;  // [Q_COUNT] totals, then decision and omega
    double *axl = tot + Q_COUNT + 8;   // [2 * n_long] sums of the long rows
    const int nnz = __ldg(&P.csr_ptr[m]);
    double *rval = axl + 2 * PD_MAX_LONG;                                  // [nnz] CSR values
    unsigned short *rcol = reinterpret_cast<unsigned short *>(rval + nnz);  // [nnz] CSR column indices
    for (int e = tid; e < nnz; e += NT) { rval[e] = __ldg(&P.csr_val[e]); rcol[e] = (unsigned short)__ldg(&P.csr_col[e]); }
    const size_t ld = P.ld;
    PdTailCols C;
    PdTailRows R;
    C.more = 0u;
#pragma unroll
    for (int k = 0; k < PD_TAIL_CPT; k++) {
        const int j = tid + k * NT;
        int beg = 0, end = 0;
        if (j < n) { beg = __ldg(&P.csc_ptr[j]); end = __ldg(&P.csc_ptr[j + 1]); }
        if (end - beg > PD_TAIL_CE) C.more |= 1u << k;
#pragma unroll
        for (int e = 0; e < PD_TAIL_CE; e++) {
            const bool has = beg + e < end;
            C.row[k][e] = has ? __ldg(&P.csc_row[beg + e]) : 0;
            C.val[k][e] = has ? __ldg(&P.csc_val[beg + e]) : 0.0;
        }
        if (j < n) {
            const size_t at = (size_t)j * ld + s;
            x[j] = Z.x[at]; xa[j] = Z.xa[at]; c[j] = Z.c[at]; lc[j] = Z.lc[at]; uc[j] = Z.uc[at];
        }
    }
#pragma unroll
    for (int k = 0; k < PD_TAIL_RPT; k++) {
        const int i = tid + k * NT;
        R.beg[k] = R.end[k] = 0; R.lidx[k] = -1;
        if (i < m) {
            R.beg[k] = __ldg(&P.csr_ptr[i]); R.end[k] = __ldg(&P.csr_ptr[i + 1]);
            if (R.end[k] - R.beg[k] > PD_LONG) {
                for (int r = 0; r < P.n_long; r++) if (__ldg(&P.long_rows[r]) == i) R.lidx[k] = r;
                R.end[k] = R.beg[k];  // the warp pass owns this row
            }
            const size_t at = (size_t)i * ld + s;
            y[i] = Z.y[at]; ya[i] = Z.ya[at]; lo[i] = Z.lo[at]; hi[i] = Z.hi[at];
        }
    }
    double omega = Z.omega[s], r0 = Z.r0[s], rprev = Z.rprev[s];
    int kbase = Z.kbase[s], it = *Z.it0, status = -1;
    const double bnorm = Z.bnorm[s], cnorm = Z.cnorm[s];
    __syncthreads();
    while (status < 0) {
        const double tau = P.eta / omega, sig = P.eta * omega;
        for (int q = 0; q < P.check; q++) {
            const double kk = (double)(it - kbase), w = (kk + 1.0) / (kk + 2.0);
#pragma unroll
            for (int k = 0; k < PD_TAIL_CPT; k++) {  // x half-step
                const int j = tid + k * NT;
                if (j < n) {
                    const double g = pd_tail_coldot(P, C, k, j, c[j], y);
                    const double xo = x[j], xn = pd_clamp(fma(-tau, g, xo), lc[j], uc[j]);
                    const double b2 = 2.0 * xn - xo;
                    xb[j] = b2;
                    x[j] = fma(w, b2, (1.0 - w) * xa[j]);
                }
            }
            __syncthreads();
            if (P.n_long > 0) { pd_tail_long_rows(P, rval, rcol, xb, xb, axl); __syncthreads(); }
#pragma unroll
            for (int k = 0; k < PD_TAIL_RPT; k++) {  // y half-step
                const int i = tid + k * NT;
                if (i < m) {
                    double ax = R.lidx[k] >= 0 ? axl[2 * R.lidx[k]] : 0.0;
                    for (int e = R.beg[k]; e < R.end[k]; e++) ax = fma(rval[e], xb[rcol[e]], ax);
                    const double yo = y[i], v = fma(-sig, ax, yo);
                    const double yn = v - pd_clamp(v, -sig * hi[i], -sig * lo[i]);
                    y[i] = fma(w, 2.0 * yn - yo, (1.0 - w) * ya[i]);
                }
            }
            __syncthreads();
            it++;
        }
        // ---- check at T(z): xn -> xb, yn -> ty -------------------------------------------------------
        double acc[Q_COUNT];
#pragma unroll
        for (int q = 0; q < Q_COUNT; q++) acc[q] = 0.0;
#pragma unroll
        for (int k = 0; k < PD_TAIL_CPT; k++) {
            const int j = tid + k * NT;
            if (j < n) {
                const double g = pd_tail_coldot(P, C, k, j, c[j], y);
                const double xn = pd_clamp(fma(-tau, g, x[j]), lc[j], uc[j]);
                xb[j] = xn;
                acc[Q_DX2] = fma(xn - x[j], xn - x[j], acc[Q_DX2]);
                acc[Q_DDX2] = fma(xn - xa[j], xn - xa[j], acc[Q_DDX2]);
                acc[Q_POBJ] = fma(c[j], xn, acc[Q_POBJ]);
                acc[Q_XN2] = fma(xn, xn, acc[Q_XN2]);
            }
        }
        __syncthreads();
        if (P.n_long > 0) { pd_tail_long_rows(P, rval, rcol, xb, x, axl); __syncthreads(); }
#pragma unroll
        for (int k = 0; k < PD_TAIL_RPT; k++) {
            const int i = tid + k * NT;
            if (i < m) {
                double axn = R.lidx[k] >= 0 ? axl[2 * R.lidx[k]] : 0.0, ax = R.lidx[k] >= 0 ? axl[2 * R.lidx[k] + 1] : 0.0;
                for (int e = R.beg[k]; e < R.end[k]; e++) { const double w2 = rval[e]; const int cc = rcol[e]; axn = fma(w2, xb[cc], axn); ax = fma(w2, x[cc], ax); }
                const double v = fma(-sig, 2.0 * axn - ax, y[i]);
                const double yn = v - pd_clamp(v, -sig * hi[i], -sig * lo[i]);
                ty[i] = yn;
                const double viol = axn - pd_clamp(axn, lo[i], hi[i]), dy = yn - y[i];
                const double yp = fmax(yn, 0.0), ym = fmax(-yn, 0.0);
                acc[Q_PR2] = fma(viol, viol, acc[Q_PR2]);
                acc[Q_DY2] = fma(dy, dy, acc[Q_DY2]);
                acc[Q_CROSS] = fma(dy, axn - ax, acc[Q_CROSS]);
                acc[Q_DDY2] = fma(yn - ya[i], yn - ya[i], acc[Q_DDY2]);
                acc[Q_DOBJR] += pd_fin(lo[i]) * yp - pd_fin(hi[i]) * ym;
                const double il = fabs(lo[i]) < PD_BIG ? 0.0 : yp, ih = fabs(hi[i]) < PD_BIG ? 0.0 : ym;
                acc[Q_DINFR] = fma(il, il, fma(ih, ih, acc[Q_DINFR]));
                acc[Q_YN2] = fma(yn, yn, acc[Q_YN2]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PD_TAIL_CPT; k++) {
            const int j = tid + k * NT;
            if (j < n) {
                const double g = pd_tail_coldot(P, C, k, j, c[j], ty);
                const double rp = fmax(g, 0.0), rm = fmax(-g, 0.0);
                const double il = fabs(lc[j]) < PD_BIG ? 0.0 : rp, iu = fabs(uc[j]) < PD_BIG ? 0.0 : rm;
                acc[Q_DRESC] = fma(il, il, fma(iu, iu, acc[Q_DRESC]));
                acc[Q_DOBJC] += pd_fin(lc[j]) * rp - pd_fin(uc[j]) * rm;
            }
        }
#pragma unroll
        for (int q = 0; q < Q_COUNT; q++) {
            double v = acc[q];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if ((tid & 31) == 0) red[(tid >> 5) * Q_COUNT + q] = v;
        }
        __syncthreads();
        if (tid < Q_COUNT) {
            double v = 0.0;
            for (int wv = 0; wv < NT / 32; wv++) v += red[wv * Q_COUNT + tid];
            tot[tid] = v;
        }
        __syncthreads();
        if (tid == 0) {  // same decision as pdhg_decide
            const double r = sqrt(fmax(tot[Q_DX2] / tau - 2.0 * tot[Q_CROSS] + tot[Q_DY2] / sig, 0.0));
            const double pobj = tot[Q_POBJ], dobj = tot[Q_DOBJR] + tot[Q_DOBJC];
            const double e_p = sqrt(tot[Q_PR2]) / (1.0 + bnorm), e_d = sqrt(tot[Q_DINFR] + tot[Q_DRESC]) / (1.0 + cnorm);
            const double e_g = fabs(pobj - dobj) / (1.0 + fabs(pobj) + fabs(dobj));
            const double rel = fmax(fmax(e_p, e_d), e_g);
            if (r0 < 0.0) r0 = r;
            const int k = it - kbase;
            double dec = 0.0;  // 0 continue, 1 restart, 2.. finished with status dec - 2
            if (rel < P.tol || it >= P.max_iter || !(rel == rel)) {
                dec = 2.0 + (rel < P.tol ? B200LP_ST_OPTIMAL : (rel == rel ? B200LP_ST_ITERLIMIT : B200LP_ST_NAN));
            } else if (r <= 0.2 * r0 || (r <= 0.8 * r0 && r > rprev) || (double)k >= 0.36 * (double)it) {
                dec = 1.0;
                omega = pd_new_omega(omega, sqrt(tot[Q_DDX2]), sqrt(tot[Q_DDY2]), sqrt(tot[Q_XN2]), sqrt(tot[Q_YN2]), e_p, e_d, e_g, bnorm, cnorm);
                r0 = -1.0;
            }
            rprev = r;
            tot[Q_COUNT] = dec; tot[Q_COUNT + 1] = omega;
        }
        __syncthreads();
        const double dec = tot[Q_COUNT];
        omega = tot[Q_COUNT + 1];
        if (dec >= 1.0) {  // restart or finished: z := anchor := T(z)
            for (int j = tid; j < n; j += NT) { const double v = xb[j]; x[j] = v; xa[j] = v; }
            for (int i = tid; i < m; i += NT) { const double v = ty[i]; y[i] = v; ya[i] = v; }
            kbase = it;
            if (dec >= 2.0) status = (int)(dec - 2.0);
        }
        __syncthreads();
    }
    for (int j = tid; j < n; j += NT) { const size_t at = (size_t)j * ld + s; Z.x[at] = x[j]; Z.xa[at] = x[j]; }
    for (int i = tid; i < m; i += NT) { const size_t at = (size_t)i * ld + s; Z.y[at] = y[i]; Z.ya[at] = y[i]; }
    if (tid == 0) {
        Z.omega[s] = omega; Z.kbase[s] = kbase; Z.r0[s] = r0; Z.rprev[s] = rprev;
        Z.done[s] = 1; Z.status[s] = status; Z.iters[s] = it;
        atomicAdd(Z.total_iters, (unsigned long long)it);
    }
}

// ---- step 5: bound snapping, un-scaling (into the xb planes), objective, node flows -------------------------
__global__ void pdhg_unscale(const PdhgDev P, const PdhgState Z, double *__restrict__ col_flow) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (s >= P.S) return;
    const size_t at = (size_t)j * P.ld + s;
    double x = 0.0;
    const int g0 = __ldg(&P.al_ptr[j]), g1 = __ldg(&P.al_ptr[j + 1]);
    for (int g = g0; g < g1; g++) {               // a folded column: the combination of the surviving columns it was replaced by
        const int root = __ldg(&P.al_col[g]);
        const size_t src = (size_t)root * P.ld + s;
        const double d = __ldg(&P.dc[root]);
        double xr = Z.x[src] * d;
        const double l = Z.lc[src] * d, u = Z.uc[src] * d;
        if (fabs(xr - l) <= 1e-9 * (1.0 + fabs(l))) xr = l;
        if (fabs(xr - u) <= 1e-9 * (1.0 + fabs(pd_fin(u)))) xr = u;
        xr = (root == j) ? xr : __ldg(&P.al_ratio[g]) * xr;
        x = (g == g0) ? xr : x + xr;
    }
    Z.xb[at] = x;
    if (col_flow) col_flow[(size_t)j * P.S + s] = x;
}

// objective = sum_j c_j x_j in the unscaled space.  Two stages with a fixed summation order (deterministic): partial sums of
// PD_OBJ_CH columns per (scenario, chunk) into Z.part, then the chunks in ascending order.  (One thread walking all the columns
// of its scenario took 1.9 ms of a 19.5 ms timestep at 8,192 scenarios x 2,672 columns.)
constexpr int PD_OBJ_CH = 128;
__global__ void pdhg_objective_partial(const PdhgDev P, const PdhgState Z) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.S) return;
    const int j0 = blockIdx.y * PD_OBJ_CH, j1 = min(P.n, j0 + PD_OBJ_CH);
    double acc = 0.0;
    for (int j = j0; j < j1; j++) acc = fma(Z.c[(size_t)j * P.ld + s] / __ldg(&P.dc[j]), Z.xb[(size_t)j * P.ld + s], acc);
    Z.part[(size_t)blockIdx.y * P.ld + s] = acc;
}
__global__ void pdhg_objective(const PdhgDev P, const PdhgState Z, double *__restrict__ objective) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.S) return;
    const int nch = (P.n + PD_OBJ_CH - 1) / PD_OBJ_CH;
    double acc = 0.0;
    for (int c = 0; c < nch; c++) acc += Z.part[(size_t)c * P.ld + s];
    Z.objective[s] = acc;
    if (objective) objective[s] = acc;
}

// node_flow[v] = sum_k flow_w[k] x[flow_col[k]]  (cython_glpk.pyx:1899-1911)
__global__ void pdhg_node_flow(const PdhgDev P, const PdhgState Z, double *__restrict__ node_flow) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (s >= P.S) return;
    double acc = 0.0;
    for (int k = __ldg(&P.flow_indptr[v]); k < __ldg(&P.flow_indptr[v + 1]); k++)
        acc = fma(__ldg(&P.flow_w[k]), Z.xb[(size_t)__ldg(&P.flow_col[k]) * P.ld + s], acc);
    node_flow[(size_t)v * P.S + s] = acc;
}

__global__ void pdhg_clear_status(const PdhgDev P, const PdhgState Z) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.ld) Z.status[s] = 0;
}

// ------------------------------------------------------------------------------------------------------
// Host side: plan (presolve + scaling), memory, the per-timestep driver.
// ------------------------------------------------------------------------------------------------------
constexpr int PD_AGG_LEN = 12;   // longest equality row the presolve substitutes away (rule 2 of PdhgPlan::build)

struct PdhgPlan {
    int m = 0, n = 0, m_r = 0;
    std::vector<int> keep, sing_ptr, sing_row;
    std::vector<double> sing_coef, dr, dc;
    // equality-row elimination: column q folded into surviving columns (x_q = sum al_ratio x_{al_col} over al_ptr[q]..al_ptr[q+1];
    // a surviving column lists itself with ratio 1), the row that said so (elim_row[q]) removed; alias_root[q] != q marks a folded
    // column; members of a surviving column = itself and every column it stands in for, with the ratio (costs are summed over them)
    std::vector<int> alias_root, elim_row, cm_ptr, cm_col, al_ptr, al_col;
    std::vector<double> cm_ratio, al_ratio;
    std::vector<int> csr_ptr, csr_col, csc_ptr, csc_row;
    std::vector<double> csr_val, csc_val;
    double norm_a = 1.0, eta = 1.0;
    std::vector<int> long_rows;

    void build(const b200lp_structure *s, const double *consts) {
        m = s->n_rows; n = s->n_cols;
        // rows with duplicates merged
        std::vector<std::vector<std::pair<int, double>>> rows(m);
        for (int i = 0; i < m; i++) {
            auto &r = rows[i];
            for (int k = s->a_indptr[i]; k < s->a_indptr[i + 1]; k++) r.push_back({s->a_col[k], s->a_val[k]});
            std::sort(r.begin(), r.end(), [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; });
            size_t w = 0;
            for (size_t k = 0; k < r.size(); k++) {
                if (w > 0 && r[w - 1].first == r[k].first) r[w - 1].second += r[k].second;
                else r[w++] = r[k];
            }
            r.resize(w);
        }
        // zero entries removed: every row is a sorted list of (column, non-zero value)
        for (int i = 0; i < m; i++) {
            auto &r = rows[i];
            size_t w = 0;
            for (size_t k = 0; k < r.size(); k++)
                if (r[k].second != 0.0) r[w++] = r[k];
            r.resize(w);
        }
        // Equality-row elimination (aggregator presolve), two rules applied until neither fires.  Same order of decisions and
        // of floating-point operations as oracle/pdhg_oracle.py.
        //  (1) doubletons: a row  a x_p + b x_q = 0  with constant zero bounds and a b < 0 - the mass balance of a link with one
        //      edge in and one edge out (cython_glpk.pyx:1364-1388) - fixes x_q = rho x_p, rho = -a / b > 0.  Column q is folded
        //      into column p in every other row, in the cost (cm_*: summed per scenario in pdhg_assemble_cols) and, through its
        //      singleton rows, in the bounds; the row disappears and x_q is rebuilt in pdhg_unscale.  Chains of links collapse.
        //  (2) confluences: a row  b x_q + sum_k a_k x_k = 0  (3..PD_AGG_LEN entries, constant zero bounds) in which x_q is the
        //      only entry of its sign - a node with one edge out and several in, or the reverse - fixes
        //      x_q = sum_k rho_k x_k, rho_k = -a_k / b > 0.  x_q >= 0 is implied by x_k >= 0, so q can go if nothing else bounds
        //      it: every singleton row on q must be one that can never bind (constant, lb <= 0, ub infinite).  q is replaced by
        //      the combination in every other row (fill-in: the downstream balance lists the upstream edges directly).
        // The basis inverse of the revised simplex shrinks quadratically with the rows removed (2,073-node network: 1,242 ->
        // 1,040 rows with rule 1, -> 561 with both; the specified config-4 network: 1,054 -> 668 -> 497).
        alias_root.resize(n); elim_row.assign(n, -1);
        for (int j = 0; j < n; j++) alias_root[j] = j;
        std::vector<std::vector<std::pair<int, double>>> comb(n);   // folded column -> its combination of surviving columns
        std::vector<char> gone(m, 0);
        auto zero_bounds = [&](int i) {
            if (s->row_kind[i] != B200LP_ROW_FLOW || s->row_lb_ref[i] >= 0 || s->row_ub_ref[i] >= 0) return false;
            return std::fabs(consts[-s->row_lb_ref[i] - 1]) < 1e-8 && std::fabs(consts[-s->row_ub_ref[i] - 1]) < 1e-8;
        };
        auto never_binds = [&](int i) {   // on x >= 0: non-negative coefficients, constant lb <= 0, no upper bound
            if (s->row_kind[i] != B200LP_ROW_FLOW || rows[i].empty() || s->row_lb_ref[i] >= 0 || s->row_ub_ref[i] >= 0) return false;
            for (auto &e : rows[i]) if (e.second < 0.0) return false;
            return consts[-s->row_lb_ref[i] - 1] < 1e-8 && consts[-s->row_ub_ref[i] - 1] >= PD_BIG;
        };
        auto count_col = [&](int col) {
            int c = 0;
            for (int k = 0; k < m; k++)
                if (!gone[k])
                    for (auto &e : rows[k]) if (e.first == col) { c++; break; }
            return c;
        };
        // list `l` (sorted by column): entry q (coefficient v) replaced by v * C, merged into the entries already there
        auto substitute = [&](std::vector<std::pair<int, double>> &l, int q, const std::vector<std::pair<int, double>> &C) {
            int hit = -1;
            for (size_t t = 0; t < l.size(); t++) if (l[t].first == q) { hit = (int)t; break; }
            if (hit < 0) return;
            const double v = l[hit].second;
            l.erase(l.begin() + hit);
            bool added = false;
            for (auto &ck : C) {
                int at = -1;
                for (size_t t = 0; t < l.size(); t++) if (l[t].first == ck.first) { at = (int)t; break; }
                const double nv = (at >= 0 ? l[at].second : 0.0) + v * ck.second;
                if (at >= 0) {
                    if (nv != 0.0) l[at].second = nv; else l.erase(l.begin() + at);
                } else if (nv != 0.0) {
                    l.push_back({ck.first, nv});
                    added = true;
                }
            }
            if (added) std::sort(l.begin(), l.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) { return x.first < y.first; });
        };
        auto eliminate = [&](int i, int q, const std::vector<std::pair<int, double>> &C) {
            for (int k = 0; k < m; k++)
                if (k != i && !gone[k]) substitute(rows[k], q, C);
            for (int j = 0; j < n; j++)
                if (!comb[j].empty()) substitute(comb[j], q, C);
            comb[q] = C;
            gone[i] = 1;
            elim_row[q] = i;
        };
        const bool rule1 = getenv("B200LP_NO_DOUBLETON") == nullptr;
        const bool rule2 = rule1 && getenv("B200LP_NO_AGGREGATE") == nullptr;
        bool changed = rule1;
        while (changed) {
            changed = false;
            std::vector<int> ccount(n, 0);
            for (int i = 0; i < m; i++)
                if (!gone[i]) for (auto &e : rows[i]) ccount[e.first]++;
            for (int i = 0; i < m; i++) {
                if (gone[i] || rows[i].size() != 2 || !zero_bounds(i)) continue;
                const int p0 = rows[i][0].first, p1 = rows[i][1].first;
                const double a = rows[i][0].second, b = rows[i][1].second;
                if (a * b >= 0.0) continue;
                const int q = (ccount[p0] < ccount[p1] || (ccount[p0] == ccount[p1] && p0 > p1)) ? p0 : p1;
                const int pcol = q == p0 ? p1 : p0;
                const double aq = q == p0 ? a : b, ap = q == p0 ? b : a;
                eliminate(i, q, {{pcol, -ap / aq}});
                ccount[q] = 0;
                ccount[pcol] = count_col(pcol);
                changed = true;
            }
            if (!rule2) continue;
            for (int i = 0; i < m; i++) {
                const int len = (int)rows[i].size();
                if (gone[i] || len < 3 || len > PD_AGG_LEN || !zero_bounds(i)) continue;
                int npos = 0, at_pos = -1, at_neg = -1;
                for (int t = 0; t < len; t++) {
                    if (rows[i][t].second > 0.0) { npos++; at_pos = t; } else at_neg = t;
                }
                const int at = npos == 1 ? at_pos : (npos == len - 1 ? at_neg : -1);
                if (at < 0) continue;
                const int q = rows[i][at].first;
                const double b = rows[i][at].second;
                bool bounded = false;
                for (int k = 0; k < m && !bounded; k++) {
                    if (k == i || gone[k] || rows[k].size() != 1 || rows[k][0].first != q) continue;
                    if (!never_binds(k)) bounded = true;
                }
                if (bounded) continue;
                std::vector<std::pair<int, double>> C;
                for (int t = 0; t < len; t++)
                    if (t != at) C.push_back({rows[i][t].first, -rows[i][t].second / b});
                eliminate(i, q, C);
                changed = true;
            }
        }
        al_ptr.assign(n + 1, 0); al_col.clear(); al_ratio.clear();
        for (int j = 0; j < n; j++) {
            if (comb[j].empty()) { al_col.push_back(j); al_ratio.push_back(1.0); }
            else {
                alias_root[j] = comb[j][0].first;   // != j: marks the column as folded
                for (auto &e : comb[j]) { al_col.push_back(e.first); al_ratio.push_back(e.second); }
            }
            al_ptr[j + 1] = (int)al_col.size();
        }
        cm_ptr.assign(n + 1, 0); cm_col.clear(); cm_ratio.clear();
        {
            std::vector<std::vector<std::pair<int, double>>> members(n);
            for (int j = 0; j < n; j++) if (comb[j].empty()) members[j].push_back({j, 1.0});
            for (int q = 0; q < n; q++) for (auto &e : comb[q]) members[e.first].push_back({q, e.second});
            for (int j = 0; j < n; j++) {
                for (auto &e : members[j]) { cm_col.push_back(e.first); cm_ratio.push_back(e.second); }
                cm_ptr[j + 1] = (int)cm_col.size();
            }
        }
        keep.clear();
        std::vector<std::vector<std::pair<int, double>>> sing(n);
        for (int i = 0; i < m; i++) {
            if (gone[i]) continue;
            int cnt = 0, last = -1;
            bool nonneg = true;
            for (size_t k = 0; k < rows[i].size(); k++) {
                if (rows[i][k].second != 0.0) { cnt++; last = (int)k; }
                if (rows[i][k].second < 0.0) nonneg = false;
            }
            if (cnt == 1) { sing[rows[i][last].first].push_back({i, rows[i][last].second}); continue; }
            if (s->row_kind[i] == B200LP_ROW_FLOW && cnt > 0 && nonneg && s->row_lb_ref[i] < 0 && s->row_ub_ref[i] < 0) {
                const double lb = consts[-s->row_lb_ref[i] - 1], ub = consts[-s->row_ub_ref[i] - 1];
                if (lb < 1e-8 && ub >= PD_BIG) continue;  // can never bind on x >= 0
            }
            keep.push_back(i);
        }
        m_r = (int)keep.size();
        sing_ptr.assign(n + 1, 0); sing_row.clear(); sing_coef.clear();
        for (int j = 0; j < n; j++) {
            for (auto &e : sing[j]) { sing_row.push_back(e.first); sing_coef.push_back(e.second); }
            sing_ptr[j + 1] = (int)sing_row.size();
        }
        // reduced CSR
        csr_ptr.assign(m_r + 1, 0); csr_col.clear(); csr_val.clear();
        for (int r = 0; r < m_r; r++) {
            for (auto &e : rows[keep[r]])
                if (e.second != 0.0) { csr_col.push_back(e.first); csr_val.push_back(e.second); }
            csr_ptr[r + 1] = (int)csr_col.size();
        }
        const int nnz = (int)csr_col.size();
        long_rows.clear();
        for (int r = 0; r < m_r; r++)
            if (csr_ptr[r + 1] - csr_ptr[r] > PD_LONG) long_rows.push_back(r);
        dr.assign(m_r, 1.0); dc.assign(n, 1.0);
        std::vector<double> rn(m_r), cn(n);
        for (int sweep = 0; sweep <= 10; sweep++) {  // 10 Ruiz sweeps (max norm) + 1 Pock-Chambolle (1-norm)
            std::fill(rn.begin(), rn.end(), 0.0); std::fill(cn.begin(), cn.end(), 0.0);
            for (int r = 0; r < m_r; r++)
                for (int k = csr_ptr[r]; k < csr_ptr[r + 1]; k++) {
                    const double a = std::fabs(csr_val[k]);
                    if (sweep < 10) { rn[r] = std::max(rn[r], a); cn[csr_col[k]] = std::max(cn[csr_col[k]], a); }
                    else { rn[r] += a; cn[csr_col[k]] += a; }
                }
            for (auto &v : rn) v = v > 0.0 ? std::sqrt(v) : 1.0;
            for (auto &v : cn) v = v > 0.0 ? std::sqrt(v) : 1.0;
            for (int r = 0; r < m_r; r++) {
                for (int k = csr_ptr[r]; k < csr_ptr[r + 1]; k++) csr_val[k] = csr_val[k] / rn[r] / cn[csr_col[k]];
                dr[r] /= rn[r];
            }
            for (int j = 0; j < n; j++) dc[j] /= cn[j];
        }
        // final values straight from the original coefficients (same as the oracle: diag(dr) A diag(dc))
        {
            int k = 0;
            for (int r = 0; r < m_r; r++)
                for (auto &e : rows[keep[r]])
                    if (e.second != 0.0) { csr_val[k] = dr[r] * e.second * dc[e.first]; k++; }
        }
        // CSC
        csc_ptr.assign(n + 1, 0); csc_row.assign(nnz, 0); csc_val.assign(nnz, 0.0);
        for (int k = 0; k < nnz; k++) csc_ptr[csr_col[k] + 1]++;
        for (int j = 0; j < n; j++) csc_ptr[j + 1] += csc_ptr[j];
        std::vector<int> pos(csc_ptr.begin(), csc_ptr.end() - 1);
        for (int r = 0; r < m_r; r++)
            for (int k = csr_ptr[r]; k < csr_ptr[r + 1]; k++) { const int p = pos[csr_col[k]]++; csc_row[p] = r; csc_val[p] = csr_val[k]; }
        // ||A||_2 by power iteration on A^T A
        std::vector<double> v(n, 1.0 / std::sqrt((double)n)), t(m_r), w(n);
        double nv = 0.0;
        for (int itp = 0; itp < 300; itp++) {
            for (int r = 0; r < m_r; r++) {
                double acc = 0.0;
                for (int k = csr_ptr[r]; k < csr_ptr[r + 1]; k++) acc += csr_val[k] * v[csr_col[k]];
                t[r] = acc;
            }
            std::fill(w.begin(), w.end(), 0.0);
            for (int r = 0; r < m_r; r++)
                for (int k = csr_ptr[r]; k < csr_ptr[r + 1]; k++) w[csr_col[k]] += csr_val[k] * t[r];
            nv = 0.0;
            for (double a : w) nv += a * a;
            nv = std::sqrt(nv);
            if (nv == 0.0) break;
            for (int j = 0; j < n; j++) v[j] = w[j] / nv;
        }
        norm_a = nv > 0.0 ? std::sqrt(nv) : 1.0;
        eta = 0.99 / norm_a;
    }
};

}  // namespace b200lp
