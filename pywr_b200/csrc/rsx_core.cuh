// rsx_core.cuh — bounded revised simplex with an explicit basis inverse, one CTA per scenario.
//
// Role (BASELINE.json north_star): the warm-started per-scenario simplex for networks whose dense dictionary
// does not fit shared memory (config 4: 2,073 nodes, edge formulation, 1,242 x 2,672 after the structural
// presolve).  It replaces CythonGLPKEdgeSolver._solve_scenario's glp_simplex call with the scenario's own
// previous basis (pywr/solvers/cython_glpk.pyx:197-232 BasisManager, :1026-1063 solve + save basis, retry
// ladder :1034-1042) on the LP assembled by the PDHG front end (same presolve, scaling and planes:
// pdhg_engine.cuh), and returns a VERTEX, which the first-order path does not.
//
// Formulation: variables v = (x, r), r = A x, i.e. [A | -I] v = 0;  m basic, n non-basic variables,
// v_B = T v_N with the dictionary T = -W N, W = B^-1.  The dictionary itself (m x n) does not fit, so only
// W (m x m doubles, row-major, in HBM: 12.3 MB per scenario for config 4) is kept per scenario and the
// pivot row / column are formed on demand from the shared sparse A:
//     row r of T:     T[r][j] = -rho . A_j  (structural j),   +rho_k  (slack of row k),   rho = W[r][:]
//     column q of T:  T[:][q] = -w,  w = W M_q,  M_q = A_q or -e_k
// Pivoting rules, tolerances and the dual-then-primal phase structure are those of the dictionary simplex
// (lp_kernel.cuh::Dict::solve, oracle/lp_oracle.c::dict_solve): largest violation leaves, ratio test on the
// (shifted) reduced costs with ties -> largest pivot -> lowest index, then Dantzig primal with bound flips.
//
// Working vectors (bounds, reduced costs, pivot row, beta, rho, w, basis heads, statuses) live in shared
// memory for the whole scenario solve; W is streamed: one full read per scenario-timestep (beta = -W g) plus,
// per pivot, one row, a few columns and the rank-1 update restricted to the non-zeros of rho and w.
//
// The file compiles for the device (DevCtx, rsx_kernels.cuh) and, unchanged, for the host with a one-thread
// context (oracle/rsx_host.cpp).  The host build is TEST infrastructure only: it lets the algorithm be checked
// against the dictionary oracle and HiGHS in a container without a GPU.  The product library contains the
// device instantiation only (no CPU fallback).
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define RSX_FN __device__ __forceinline__
#define RSX_HD __host__ __device__ inline
#else
#define RSX_FN static inline
#define RSX_HD static inline
#endif

// read-only data shared by all scenarios (the sparse matrix): ld.global.nc on the device, so that the compiler may hoist these
// loads above the stores of the surrounding loops (it cannot prove that the work vectors do not alias them)
#ifdef __CUDA_ARCH__
#define RSX_RO(p) __ldg(p)
#else
#define RSX_RO(p) (*(p))
#endif

namespace rsx {

constexpr double TOL_PRIMAL = 1e-9, TOL_DUAL = 1e-9, TOL_PIVOT = 1e-9, TIE_REL = 1e-9, TIE_ABS = 1e-12;
constexpr double BIG = 1e300;        // |bound| >= BIG is infinite (same convention as pdhg_engine.cuh)
constexpr double DROP = 1e-14;       // |w_i| below this is treated as zero in the rank-1 update
enum : int { V_BASIC = 1, V_NL = 2, V_NU = 3, V_NF = 4, V_NS = 5 };
// B200LP_ST_* codes (include/b200lp.h) + one internal code
enum : int { R_OPTIMAL = 0, R_INFEASIBLE = 1, R_UNBOUNDED = 2, R_ITERLIMIT = 3, R_NAN = 4, R_SINGULAR = 5, R_NUMERIC = 100 };

struct Lp {  // shared by all scenarios
    int m, n, ldw;  // ldw: row stride of W (m rounded up to even)
    const int *csr_ptr, *csr_col;
    const double *csr_val;
    const int *csc_ptr, *csc_row;
    const double *csc_val;
    int cost_dynamic;  // costs change between timesteps: reduced costs are rebuilt every solve
    // rows with more than `long_len` entries (aggregated nodes: one licence row spans every abstraction): their sums are
    // taken by a whole warp, everything else by one thread per row (a 200-entry row walked by one thread is 200 dependent
    // L2 round trips: 160 us of a 470 us scenario-timestep)
    int n_long, long_len;
    const int *long_rows;
    // the reduced costs are carried from pivot to pivot and from timestep to timestep; after this many pivots of a scenario
    // they are rebuilt from W and the costs (bounds the round-off a 30-year run could accumulate); 0: never
    int refresh_pivots;
};

struct Scen {  // one scenario: persistent state + this timestep's planes (element stride ld)
    double *W;            // [m][ldw]
    int *head;            // [m]
    unsigned char *stat;  // [n+m]
    double *d;            // [n+m]
    int *info;            // [4]: valid, pivots since the last refactorisation, refactorisations, unused
    const double *c, *lc, *uc, *lo, *hi;
    double *x;
    size_t ld;
    // device: the scenario's bounds gathered into one contiguous vector [LO(n+m) | HI(n+m)] and its x written contiguously
    // (rsx_gather_bounds / rsx_scatter_x transpose them from / to the planes with coalesced accesses on both sides): 14 k
    // plane elements per scenario-timestep read 64 KB apart cost the stream of W a third of its bandwidth
    const double *bt;  // nullptr: read the planes
    size_t ldx;        // element stride of x
};

struct Work {  // block-shared
    double *LO, *HI, *d, *dw, *alpha;  // [n+m]
    double *rho, *w, *beta, *g;        // [m]
    int *head, *list_k, *list_i;       // [m]
    unsigned char *stat;               // [n+m]
    int *cnt;                          // [8]
};

// Placement of the working vectors.  mode 0: everything in the block's fast memory (shared memory on the device).
// mode 1: the reduced costs (true / working) and the two index lists live in a per-block scratch area in
// global memory (L2-resident: 73 KB per block for config 4) — they are touched once per timestep and per pivot, with
// coalesced accesses — which halves the shared-memory footprint: two blocks per SM, so that one scenario's serial phases
// overlap another's stream of W, and LPs twice as large still fit.
// g (right-hand side N x_N at the start of a solve, row activities at its end, y in compute_d) and w (the pivot column,
// live only inside a pivot) are never live together: they share one vector.
RSX_HD size_t work_bytes(int m, int n, int mode) {
    const size_t nv = (size_t)n + m;
    const size_t fast_nv = mode == 0 ? 5 : 3, lists = mode == 0 ? 3 : 1;
    return 8 * (fast_nv * nv + 3 * (size_t)m) + 4 * (lists * (size_t)m) + ((nv + 15) & ~(size_t)15) + 64;
}
RSX_HD size_t scratch_bytes(int m, int n, int mode) {
    const size_t nv = (size_t)n + m;
    return mode == 0 ? 0 : 8 * (2 * nv) + 4 * (2 * (size_t)m) + 64;
}
RSX_FN void work_carve(Work &K, unsigned char *base, unsigned char *scratch, int m, int n, int mode) {
    const size_t nv = (size_t)n + m;
    double *p = reinterpret_cast<double *>(base);
    K.LO = p; p += nv; K.HI = p; p += nv;
    if (mode == 0) { K.d = p; p += nv; K.dw = p; p += nv; }
    K.alpha = p; p += nv;  // the pivot row is read three to four times per pivot: always in fast memory
    K.rho = p; p += m; K.w = p; K.g = p; p += m; K.beta = p; p += m;
    int *q = reinterpret_cast<int *>(p);
    K.head = q; q += m;
    if (mode == 0) { K.list_k = q; q += m; K.list_i = q; q += m; }
    K.cnt = q; q += 8;
    K.stat = reinterpret_cast<unsigned char *>(q);
    if (mode != 0) {
        double *sp = reinterpret_cast<double *>(scratch);
        K.d = sp; sp += nv; K.dw = sp; sp += nv;
        int *sq = reinterpret_cast<int *>(sp);
        K.list_k = sq; sq += m; K.list_i = sq;
    }
}

// value of a non-basic variable
RSX_FN double nb_value(const Work &K, int v) {
    const int s = K.stat[v];
    return s == V_NU ? K.HI[v] : (s == V_NF ? 0.0 : K.LO[v]);
}

template <class CX>
RSX_FN void load_bounds(CX &cx, const Lp &P, const Scen &S, Work &K) {
    const int n = P.n, nv = P.n + P.m;
    for (int v = cx.tid; v < nv; v += cx.nt) {
        double lo, hi;
        if (S.bt != nullptr) { lo = S.bt[v]; hi = S.bt[nv + v]; }
        else if (v < n) { lo = S.lc[(size_t)v * S.ld]; hi = S.uc[(size_t)v * S.ld]; }
        else { lo = S.lo[(size_t)(v - n) * S.ld]; hi = S.hi[(size_t)(v - n) * S.ld]; }
        K.LO[v] = lo <= -BIG ? -INFINITY : lo;
        K.HI[v] = hi >= BIG ? INFINITY : hi;
    }
    cx.sync();
}

// slack basis: W = -I, every structural non-basic at its lower bound, d = c
template <class CX>
RSX_FN void cold_init(CX &cx, const Lp &P, const Scen &S, Work &K) {
    const int m = P.m, n = P.n, nv = n + m;
    for (int v = cx.tid; v < nv; v += cx.nt) {
        K.stat[v] = v < n ? V_NL : V_BASIC;
        K.d[v] = v < n ? S.c[(size_t)v * S.ld] : 0.0;
    }
    for (int i = cx.tid; i < m; i += cx.nt) K.head[i] = n + i;
    const size_t tot = (size_t)m * P.ldw;
    for (size_t e = cx.tid; e < tot; e += cx.nt) S.W[e] = 0.0;
    cx.sync();
    for (int i = cx.tid; i < m; i += cx.nt) S.W[(size_t)i * P.ldw + i] = -1.0;
    cx.sync();
}

template <class CX>
RSX_FN void load_rho(CX &cx, const Lp &P, const Scen &S, Work &K, int r) {
    const double *row = S.W + (size_t)r * P.ldw;
    for (int k = cx.tid; k < P.m; k += cx.nt) K.rho[k] = row[k];
    cx.sync();
}

// alpha[v] = T[r][v] for every non-basic v (0 for basic ones); needs rho
template <class CX>
RSX_FN void compute_alpha(CX &cx, const Lp &P, Work &K) {
    const int n = P.n, nv = P.n + P.m;
    for (int v = cx.tid; v < nv; v += cx.nt) {
        double a = 0.0;
        if (K.stat[v] != V_BASIC) {
            if (v < n) {
                const int e1 = RSX_RO(&P.csc_ptr[v + 1]);
                for (int e = RSX_RO(&P.csc_ptr[v]); e < e1; e++) a = fma(K.rho[RSX_RO(&P.csc_row[e])], RSX_RO(&P.csc_val[e]), a);
                a = -a;
            } else {
                a = K.rho[v - n];
            }
        }
        K.alpha[v] = a;
    }
    cx.sync();
}

// w = W M_q: a gather of the few columns of W that M_q touches (W is row-major: every element is its own DRAM sector).
// Each thread takes four rows at a time and up to four column entries, so that up to 16 independent loads are in flight per
// thread instead of one dependent DRAM round trip after the other (that was 14 k cycles per pivot).
template <class CX>
RSX_FN void compute_w(CX &cx, const Lp &P, const Scen &S, Work &K, int q) {
    const int m = P.m, n = P.n;
    int ke[4] = {0, 0, 0, 0};
    double ae[4] = {0.0, 0.0, 0.0, 0.0};
    int e0 = 0, e1 = 0;
    if (q < n) {
        e0 = RSX_RO(&P.csc_ptr[q]); e1 = RSX_RO(&P.csc_ptr[q + 1]);
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (e0 + t < e1) { ke[t] = RSX_RO(&P.csc_row[e0 + t]); ae[t] = RSX_RO(&P.csc_val[e0 + t]); }
    } else {
        ke[0] = q - n; ae[0] = -1.0; e0 = 0; e1 = 1;
    }
    const int ne = e1 - e0 < 4 ? e1 - e0 : 4;
    for (int base = cx.tid; base < m; base += 4 * cx.nt) {
        double x[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + u * cx.nt;
            const double *row = S.W + (size_t)(i < m ? i : 0) * P.ldw;
#pragma unroll
            for (int t = 0; t < 4; t++) x[u][t] = (i < m && t < ne) ? row[ke[t]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + u * cx.nt;
            if (i >= m) continue;
            double a = 0.0;
#pragma unroll
            for (int t = 0; t < 4; t++) a = fma(x[u][t], ae[t], a);
            if (q < n)
                for (int e = e0 + 4; e < e1; e++) a = fma(S.W[(size_t)i * P.ldw + RSX_RO(&P.csc_row[e])], RSX_RO(&P.csc_val[e]), a);
            K.w[i] = a;
        }
    }
    cx.sync();
}

// rank-1 update of W for the exchange in basic position r (rho = old row r, w = W M_q), restricted to the
// non-zeros of rho and w.  On return rho holds the new row r.
template <class CX>
RSX_FN void update_W(CX &cx, const Lp &P, const Scen &S, Work &K, int r) {
    const int m = P.m;
    if (cx.leader()) { K.cnt[0] = 0; K.cnt[1] = 0; }
    cx.sync();
    const double inv = 1.0 / K.w[r];
    for (int k = cx.tid; k < m; k += cx.nt) {
        const double v = K.rho[k];
        if (v != 0.0) K.list_k[cx.atomic_inc(&K.cnt[0])] = k;
        K.rho[k] = v * inv;
    }
    for (int i = cx.tid; i < m; i += cx.nt)
        if (i != r && fabs(K.w[i]) > DROP) K.list_i[cx.atomic_inc(&K.cnt[1])] = i;
    cx.sync();
    const int nk = K.cnt[0], ni = K.cnt[1];
    const bool sparse = nk * 4 < m;
    double *rowr = S.W + (size_t)r * P.ldw;
    if (sparse) { for (int t = cx.tid; t < nk; t += cx.nt) { const int k = K.list_k[t]; rowr[k] = K.rho[k]; } }
    else { for (int k = cx.tid; k < m; k += cx.nt) rowr[k] = K.rho[k]; }
    // two rows per warp at a time, four segments of each in flight: the update is a chain of read-modify-write round trips to
    // L2 / HBM, and with one row per warp a pivot spent 12-15 k cycles here
    for (int li = 2 * cx.warp; li < ni; li += 2 * cx.nw) {
        const int i0 = K.list_i[li];
        const bool two = li + 1 < ni;
        const int i1 = two ? K.list_i[li + 1] : i0;
        const double w0 = -K.w[i0], w1 = -K.w[i1];
        double *r0 = S.W + (size_t)i0 * P.ldw, *r1 = S.W + (size_t)i1 * P.ldw;
        constexpr int nl = CX::nl;
        if (sparse) {
            for (int t = cx.lane; t < nk; t += nl) {
                const int k = K.list_k[t];
                const double rk = K.rho[k], x0 = r0[k], x1 = r1[k];
                r0[k] = fma(w0, rk, x0);
                if (two) r1[k] = fma(w1, rk, x1);
            }
        } else {
            int k = cx.lane;
            for (; k + 3 * nl < m; k += 4 * nl) {
                const double x0 = r0[k], x1 = r0[k + nl], x2 = r0[k + 2 * nl], x3 = r0[k + 3 * nl];
                const double y0 = r1[k], y1 = r1[k + nl], y2 = r1[k + 2 * nl], y3 = r1[k + 3 * nl];
                const double g0 = K.rho[k], g1 = K.rho[k + nl], g2 = K.rho[k + 2 * nl], g3 = K.rho[k + 3 * nl];
                r0[k] = fma(w0, g0, x0); r0[k + nl] = fma(w0, g1, x1); r0[k + 2 * nl] = fma(w0, g2, x2); r0[k + 3 * nl] = fma(w0, g3, x3);
                if (two) { r1[k] = fma(w1, g0, y0); r1[k + nl] = fma(w1, g1, y1); r1[k + 2 * nl] = fma(w1, g2, y2); r1[k + 3 * nl] = fma(w1, g3, y3); }
            }
            for (; k < m; k += nl) {
                const double g = K.rho[k], x = r0[k], y = r1[k];
                r0[k] = fma(w0, g, x);
                if (two) r1[k] = fma(w1, g, y);
            }
        }
    }
    cx.sync();
}

// reduced costs from scratch: y = W^T c_B, d_j = c_j - y . A_j (structural), y_k (slack of row k)
template <class CX>
RSX_FN void compute_d(CX &cx, const Lp &P, const Scen &S, Work &K) {
    const int m = P.m, n = P.n, nv = n + m;
    for (int i = cx.tid; i < m; i += cx.nt) { const int v = K.head[i]; K.rho[i] = v < n ? S.c[(size_t)v * S.ld] : 0.0; }
    cx.sync();
    for (int k = cx.tid; k < m; k += cx.nt) {
        double acc = 0.0;
        for (int i = 0; i < m; i++) {
            const double cb = K.rho[i];
            if (cb != 0.0) acc = fma(cb, S.W[(size_t)i * P.ldw + k], acc);
        }
        K.g[k] = acc;
    }
    cx.sync();
    for (int v = cx.tid; v < nv; v += cx.nt) {
        double dv = 0.0;
        if (K.stat[v] != V_BASIC) {
            if (v < n) {
                double a = 0.0;
                for (int e = P.csc_ptr[v]; e < P.csc_ptr[v + 1]; e++) a = fma(K.g[P.csc_row[e]], P.csc_val[e], a);
                dv = S.c[(size_t)v * S.ld] - a;
            } else {
                dv = K.g[v - n];
            }
        }
        K.d[v] = dv;
    }
    cx.sync();
}

// status of every non-basic variable for the current bounds (lp_kernel.cuh::place_nonbasics_lane);
// returns 1 if some reduced cost is dual infeasible for the bound the variable has to sit on
template <class CX>
RSX_FN int place_nonbasics(CX &cx, const Lp &P, Work &K) {
    const int nv = P.n + P.m;
    int infeas = 0;
    for (int v = cx.tid; v < nv; v += cx.nt) {
        const int s0 = K.stat[v];
        if (s0 == V_BASIC) continue;
        const double lo = K.LO[v], hi = K.HI[v], dj = K.d[v];
        const bool lo_fin = lo > -INFINITY, hi_fin = hi < INFINITY, fx = lo == hi;
        const bool dpos = dj > TOL_DUAL, dneg = dj < -TOL_DUAL;
        const int s_box = dpos ? V_NL : (dneg ? V_NU : ((s0 == V_NL || s0 == V_NU) ? s0 : V_NL));
        int s = lo_fin ? (hi_fin ? s_box : V_NL) : (hi_fin ? V_NU : V_NF);
        if (fx) s = V_NS;
        const bool bad = lo_fin ? (!hi_fin && dneg) : (hi_fin ? dpos : (dpos || dneg));
        if (!fx && bad) infeas = 1;
        K.stat[v] = (unsigned char)s;
    }
    infeas = cx.bor(infeas);
    return infeas;
}

// beta = -W g,  g = N x_N  (one full read of W)
template <class CX>
RSX_FN void compute_beta(CX &cx, const Lp &P, const Scen &S, Work &K) {
    const int m = P.m, n = P.n;
    for (int i = cx.tid; i < m; i += cx.nt) {
        const int e0 = RSX_RO(&P.csr_ptr[i]), e1 = RSX_RO(&P.csr_ptr[i + 1]);
        if (e1 - e0 > P.long_len) continue;
        double acc = 0.0;
        for (int e = e0; e < e1; e++) {
            const int j = RSX_RO(&P.csr_col[e]);
            if (K.stat[j] != V_BASIC) acc = fma(RSX_RO(&P.csr_val[e]), nb_value(K, j), acc);
        }
        if (K.stat[n + i] != V_BASIC) acc -= nb_value(K, n + i);
        K.g[i] = acc;
    }
    for (int li = cx.warp; li < P.n_long; li += cx.nw) {
        const int i = P.long_rows[li];
        double acc = 0.0;
        for (int e = P.csr_ptr[i] + cx.lane; e < P.csr_ptr[i + 1]; e += cx.nl) {
            const int j = P.csr_col[e];
            if (K.stat[j] != V_BASIC) acc = fma(P.csr_val[e], nb_value(K, j), acc);
        }
        acc = cx.warp_sum(acc);
        if (K.stat[n + i] != V_BASIC) acc -= nb_value(K, n + i);
        if (cx.lane == 0) K.g[i] = acc;
    }
    cx.sync();
    cx.mark(4);
    // two rows per warp and four independent 256-byte segments per row in flight: the pass is a pure HBM stream
    for (int i = 2 * cx.warp; i < m; i += 2 * cx.nw) {
        const double *r0 = S.W + (size_t)i * P.ldw;
        const bool two = i + 1 < m;
        const double *r1 = two ? r0 + P.ldw : r0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        int k = cx.lane;
        constexpr int nl = CX::nl;
        // 4 x unrolled: 32 independent 256-byte loads per warp in flight (128 KB per SM); with 8 the stream reached 5.1 of the
        // 7.1 TB/s the same loop sustains when the compiler unrolls it by itself (benchmarks/tools/stream_probe.cu)
#pragma unroll 4
        for (; k + 3 * nl < m; k += 4 * nl) {
            const double x0 = r0[k], x1 = r0[k + nl], x2 = r0[k + 2 * nl], x3 = r0[k + 3 * nl];
            const double y0 = r1[k], y1 = r1[k + nl], y2 = r1[k + 2 * nl], y3 = r1[k + 3 * nl];
            const double g0 = K.g[k], g1 = K.g[k + nl], g2 = K.g[k + 2 * nl], g3 = K.g[k + 3 * nl];
            a0 = fma(x0, g0, a0); a1 = fma(x1, g1, a1); a2 = fma(x2, g2, a2); a3 = fma(x3, g3, a3);
            b0 = fma(y0, g0, b0); b1 = fma(y1, g1, b1); b2 = fma(y2, g2, b2); b3 = fma(y3, g3, b3);
        }
        for (; k < m; k += nl) { a0 = fma(r0[k], K.g[k], a0); b0 = fma(r1[k], K.g[k], b0); }
        const double sa = cx.warp_sum((a0 + a1) + (a2 + a3)), sb = cx.warp_sum((b0 + b1) + (b2 + b3));
        if (cx.lane == 0) { K.beta[i] = -sa; if (two) K.beta[i + 1] = -sb; }
    }
    cx.sync();
}

template <class CX>
RSX_FN void init_dw(CX &cx, const Lp &P, Work &K, int shifted) {
    const int nv = P.n + P.m;
    for (int base = cx.tid; base < nv; base += 4 * cx.nt) {
        double w[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int v = base + u * cx.nt; w[u] = v < nv ? K.d[v] : 0.0; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int v = base + u * cx.nt;
            if (v >= nv) continue;
            if (shifted) {
                const int s = K.stat[v];
                if ((s == V_NL && w[u] < -TOL_DUAL) || (s == V_NU && w[u] > TOL_DUAL) || (s == V_NF && fabs(w[u]) > TOL_DUAL)) w[u] = 0.0;
            }
            K.dw[v] = w[u];
        }
    }
    cx.sync();
}

// exchange basic position r (variable p) with non-basic q; rho, alpha (row r) and w (column q) are current
template <class CX>
RSX_FN void do_pivot(CX &cx, const Lp &P, const Scen &S, Work &K, int r, int q, int to_upper, bool use_dw) {
    const int m = P.m, nv = P.n + P.m;
    const int p = K.head[r];
    const double plo = K.LO[p], phi = K.HI[p];
    const double a_rq = -K.w[r];
    const double target = to_upper ? phi : plo;
    const double theta = (target - K.beta[r]) / a_rq;  // change of v_q
    const double xq_old = nb_value(K, q);
    const double rd = K.d[q] / a_rq, rw = use_dw ? K.dw[q] / a_rq : 0.0;
    cx.sync();
    cx.mark(9);   // leaving / entering selection, pivot row and column
    // four elements per thread at a time: all loads of a batch are issued before its stores (d / dw may live in the L2 scratch,
    // where one dependent round trip per element made this loop 16 k cycles per pivot)
    for (int base = cx.tid; base < nv; base += 4 * cx.nt) {
        double dv[4], wv[4], av[4];
        bool on[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int v = base + u * cx.nt;
            on[u] = v < nv && K.stat[v < nv ? v : 0] != V_BASIC && v != q;
            av[u] = on[u] ? K.alpha[v] : 0.0;
            on[u] = on[u] && av[u] != 0.0;   // the pivot row is sparse: most reduced costs do not move (and are not fetched)
            dv[u] = on[u] ? K.d[v] : 0.0;
            wv[u] = (on[u] && use_dw) ? K.dw[v] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int v = base + u * cx.nt;
            if (on[u]) {
                K.d[v] = fma(-rd, av[u], dv[u]);
                if (use_dw) K.dw[v] = fma(-rw, av[u], wv[u]);
            }
        }
    }
    for (int i = cx.tid; i < m; i += cx.nt)
        if (i != r) K.beta[i] = fma(-theta, K.w[i], K.beta[i]);
    cx.sync();
    if (cx.leader()) {
        K.d[p] = rd; K.d[q] = 0.0;
        if (use_dw) { K.dw[p] = rw; K.dw[q] = 0.0; }
        K.beta[r] = xq_old + theta;
        K.head[r] = q;
        K.stat[q] = V_BASIC;
        K.stat[p] = (unsigned char)((plo == phi) ? V_NS : (to_upper ? V_NU : V_NL));
    }
    cx.mark(10);  // reduced costs and basic values
    update_W(cx, P, S, K, r);  // syncs
    cx.mark(11);  // rank-1 update of W
}

// the simplex proper (Dict::solve): dual phase on the (shifted) reduced costs, then primal clean-up
template <class CX>
RSX_FN int solve(CX &cx, const Lp &P, const Scen &S, Work &K, int &pivots) {
    const int m = P.m, nv = P.n + P.m;
    const int cap_bland = 20 * nv + 50, cap_fail = 200 * nv + 1000;
    int shifted = place_nonbasics(cx, P, K);
    cx.sync();
    cx.mark(3);
    compute_beta(cx, P, S, K);
    cx.mark(5);
    bool dw_ready = false;
    int phase = 0;
    for (int it = 0;; it++) {
        if (it > cap_fail) return R_ITERLIMIT;
        const bool bland = it > cap_bland;
        int r, q, to_upper;
        if (phase == 0) {
            double worst = -1.0; int bi = -1, bd = 0;
            for (int i = cx.tid; i < m; i += cx.nt) {
                const int v = K.head[i];
                const double b = K.beta[i], lo = K.LO[v], hi = K.HI[v];
                double viol = 0.0; int dd = 0;
                if (b < lo - TOL_PRIMAL * (1.0 + fabs(lo))) { viol = lo - b; dd = 1; }
                else if (b > hi + TOL_PRIMAL * (1.0 + fabs(hi))) { viol = b - hi; dd = -1; }
                if (b != b) { viol = INFINITY; dd = 2; }
                if (dd != 0) {
                    if (bland) viol = 1.0;
                    if (bi < 0 || viol > worst) { worst = viol; bi = i; bd = dd; }
                }
            }
            cx.argmax(worst, bi, bd);
            if (bi < 0) {
                if (!shifted) return R_OPTIMAL;
                phase = 1; shifted = 0;
                continue;
            }
            if (bd == 2) return R_NUMERIC;
            // working reduced costs: a copy with the dual infeasibilities shifted away - only when there are any; otherwise (the
            // usual warm start) the true reduced costs serve as they are and nothing is copied or updated twice
            if (shifted && !dw_ready) { init_dw(cx, P, K, shifted); dw_ready = true; }
            const double *rcw = shifted ? K.dw : K.d;
            r = bi;
            const double dir = (double)bd;
            cx.mark(6);
            load_rho(cx, P, S, K, r);
            cx.mark(12);
            compute_alpha(cx, P, K);
            cx.mark(13);
            // both passes of the ratio test visit the eligible non-basic variables in ascending order, four per thread at a time
            // with the working reduced costs of a batch loaded together (they may live in the L2 scratch: one dependent round
            // trip per eligible variable made the two passes the largest part of a pivot)
            auto scan = [&](auto &&visit) {
                for (int base = cx.tid; base < nv; base += 4 * cx.nt) {
                    double a[4], dwv[4];
                    bool ok[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int v = base + u * cx.nt;
                        const int s = v < nv ? K.stat[v] : V_BASIC;
                        a[u] = v < nv ? K.alpha[v] * dir : 0.0;
                        ok[u] = (s == V_NL && a[u] > TOL_PIVOT) || (s == V_NU && a[u] < -TOL_PIVOT) || (s == V_NF && fabs(a[u]) > TOL_PIVOT);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) dwv[u] = ok[u] ? rcw[base + u * cx.nt] : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (ok[u]) visit(base + u * cx.nt, a[u], dwv[u]);
                }
            };
            double rmin = INFINITY;
            double amax = -1.0; int aux = 0;
            q = -1;
            // up to RC x 4 variables per thread: the ratios of the first pass stay in registers for the second (a double division
            // is a ~30-instruction subroutine: recomputing them was 13 % of the kernel's instructions, and the second pass would
            // fetch the working reduced costs again)
            constexpr int RC = 2;
            if (nv <= RC * 4 * cx.nt) {
                double rc[RC][4];
#pragma unroll
                for (int bb = 0; bb < RC; bb++) {
                    const int base = cx.tid + bb * 4 * cx.nt;
                    double a[4], dwv[4];
                    bool ok[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int v = base + u * cx.nt;
                        const int s = v < nv ? K.stat[v] : V_BASIC;
                        a[u] = v < nv ? K.alpha[v] * dir : 0.0;
                        ok[u] = (s == V_NL && a[u] > TOL_PIVOT) || (s == V_NU && a[u] < -TOL_PIVOT) || (s == V_NF && fabs(a[u]) > TOL_PIVOT);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) dwv[u] = ok[u] ? rcw[base + u * cx.nt] : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        rc[bb][u] = ok[u] ? fabs(dwv[u]) / fabs(a[u]) : INFINITY;
                        rmin = fmin(rmin, rc[bb][u]);
                    }
                }
                rmin = cx.min(rmin);
                if (rmin == INFINITY) return R_INFEASIBLE;
                const double thr = rmin * (1.0 + TIE_REL) + TIE_ABS;
#pragma unroll
                for (int bb = 0; bb < RC; bb++)
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (rc[bb][u] > thr) continue;
                        const int v = cx.tid + (bb * 4 + u) * cx.nt;
                        const double key = bland ? 1.0 : fabs(K.alpha[v]);
                        if (q < 0 || key > amax) { amax = key; q = v; }
                    }
            } else {
                scan([&](int, double a, double dwv) { rmin = fmin(rmin, fabs(dwv) / fabs(a)); });
                rmin = cx.min(rmin);
                if (rmin == INFINITY) return R_INFEASIBLE;
                const double thr = rmin * (1.0 + TIE_REL) + TIE_ABS;
                scan([&](int v, double a, double dwv) {
                    if (fabs(dwv) / fabs(a) > thr) return;
                    const double key = bland ? 1.0 : fabs(a);
                    if (q < 0 || key > amax) { amax = key; q = v; }
                });
            }
            cx.argmax(amax, q, aux);
            cx.mark(14);
            to_upper = bd < 0;  // the leaving variable goes to the bound it violated
            compute_w(cx, P, S, K, q);
            if (!(fabs(K.alpha[q] + K.w[r]) <= 1e-6 * (1.0 + fabs(K.alpha[q])))) return R_NUMERIC;  // W has drifted
        } else {
            double worst = -1.0; int dirn = 0;
            q = -1;
            for (int v = cx.tid; v < nv; v += cx.nt) {
                const int s = K.stat[v];
                if (s == V_BASIC || s == V_NS) continue;
                const double dj = K.d[v];
                int dd = 0;
                if ((s == V_NL || s == V_NF) && dj < -TOL_DUAL) dd = 1;
                else if ((s == V_NU || s == V_NF) && dj > TOL_DUAL) dd = -1;
                if (dd != 0) {
                    const double key = bland ? 1.0 : fabs(dj);
                    if (q < 0 || key > worst) { worst = key; q = v; dirn = dd; }
                }
            }
            cx.argmax(worst, q, dirn);
            if (q < 0) return R_OPTIMAL;
            const double dir = (double)dirn;
            const double inlo = K.LO[q], inhi = K.HI[q];
            double tmin = inhi - inlo;
            if (!(tmin < INFINITY)) tmin = INFINITY;
            compute_w(cx, P, S, K, q);
            double step_min = INFINITY;
            for (int i = cx.tid; i < m; i += cx.nt) {
                const double a = -K.w[i] * dir;
                const int v = K.head[i];
                double step = INFINITY;
                if (a > TOL_PIVOT) { if (K.HI[v] < INFINITY) step = (K.HI[v] - K.beta[i]) / a; }
                else if (a < -TOL_PIVOT) { if (K.LO[v] > -INFINITY) step = (K.LO[v] - K.beta[i]) / a; }
                if (step < 0.0) step = 0.0;
                step_min = fmin(step_min, step);
            }
            step_min = cx.min(step_min);
            if (step_min == INFINITY && tmin == INFINITY) return R_UNBOUNDED;
            if (tmin <= step_min) {  // bound flip of the entering variable
                const double delta = dir * tmin;
                for (int i = cx.tid; i < m; i += cx.nt) K.beta[i] = fma(-delta, K.w[i], K.beta[i]);
                if (cx.leader()) K.stat[q] = (unsigned char)(dirn > 0 ? V_NU : V_NL);
                cx.sync();
                continue;
            }
            const double thr = step_min * (1.0 + TIE_REL) + TIE_ABS;
            double amax = -1.0; int bi = -1, up = 0;
            for (int i = cx.tid; i < m; i += cx.nt) {
                const double a = -K.w[i] * dir;
                const int v = K.head[i];
                double step = INFINITY; int u = 0;
                if (a > TOL_PIVOT) { if (K.HI[v] < INFINITY) { step = (K.HI[v] - K.beta[i]) / a; u = 1; } }
                else if (a < -TOL_PIVOT) { if (K.LO[v] > -INFINITY) step = (K.LO[v] - K.beta[i]) / a; }
                if (step < 0.0) step = 0.0;
                if (step > thr) continue;
                const double key = bland ? 1.0 : fabs(a);
                if (bi < 0 || key > amax) { amax = key; bi = i; up = u; }
            }
            cx.argmax(amax, bi, up);
            if (bi < 0) return R_NUMERIC;
            r = bi; to_upper = up;
            load_rho(cx, P, S, K, r);
            compute_alpha(cx, P, K);
        }
        do_pivot(cx, P, S, K, r, q, to_upper, phase == 0 && shifted);
        pivots++;
    }
}

// Rebuild W for the basis recorded in stat[] by pivoting the basic structurals into the slack basis
// (Dict::refactor).  Returns 0, or -1 if the basis is singular.
template <class CX>
RSX_FN int refactor(CX &cx, const Lp &P, const Scen &S, Work &K) {
    const int m = P.m, n = P.n;
    for (int i = cx.tid; i < m; i += cx.nt) K.head[i] = n + i;
    const size_t tot = (size_t)m * P.ldw;
    for (size_t e = cx.tid; e < tot; e += cx.nt) S.W[e] = 0.0;
    cx.sync();
    for (int i = cx.tid; i < m; i += cx.nt) S.W[(size_t)i * P.ldw + i] = -1.0;
    cx.sync();
    for (int q = 0; q < n; q++) {
        if (K.stat[q] != V_BASIC) continue;
        compute_w(cx, P, S, K, q);
        double best = TOL_PIVOT; int bi = -1, aux = 0;
        for (int i = cx.tid; i < m; i += cx.nt) {
            const int v = K.head[i];
            if (v < n || K.stat[v] == V_BASIC) continue;  // already a structural, or a slack that stays basic
            const double a = fabs(K.w[i]);
            if (a > best) { best = a; bi = i; }
        }
        if (bi < 0) best = -1.0;
        cx.argmax(best, bi, aux);
        if (bi < 0) return -1;
        load_rho(cx, P, S, K, bi);
        update_W(cx, P, S, K, bi);
        if (cx.leader()) K.head[bi] = q;
        cx.sync();
    }
    return 0;
}

// x of the structurals into the plane, row activities into g; returns the largest relative residual |A x - r|
template <class CX>
RSX_FN double write_solution(CX &cx, const Lp &P, const Scen &S, Work &K) {
    const int m = P.m, n = P.n, nv = n + m;
    for (int v = cx.tid; v < nv; v += cx.nt) {
        if (K.stat[v] == V_BASIC) continue;
        const double val = nb_value(K, v);
        if (v < n) S.x[(size_t)v * S.ldx] = val; else K.g[v - n] = val;
    }
    for (int i = cx.tid; i < m; i += cx.nt) {
        const int v = K.head[i];
        if (v < n) S.x[(size_t)v * S.ldx] = K.beta[i]; else K.g[v - n] = K.beta[i];
    }
    cx.sync();
    double worst = -1.0; int bi = -1, aux = 0;
    for (int i = cx.tid; i < m; i += cx.nt) {
        const int e0 = RSX_RO(&P.csr_ptr[i]), e1 = RSX_RO(&P.csr_ptr[i + 1]);
        if (e1 - e0 > P.long_len) continue;
        double acc = 0.0;
        for (int e = e0; e < e1; e++) acc = fma(RSX_RO(&P.csr_val[e]), S.x[(size_t)RSX_RO(&P.csr_col[e]) * S.ldx], acc);
        const double gi = K.g[i];
        double res = fabs(acc - gi) / (1.0 + fabs(gi));
        if (res != res) res = INFINITY;
        if (bi < 0 || res > worst) { worst = res; bi = i; }
    }
    for (int li = cx.warp; li < P.n_long; li += cx.nw) {
        const int i = P.long_rows[li];
        double acc = 0.0;
        for (int e = P.csr_ptr[i] + cx.lane; e < P.csr_ptr[i + 1]; e += cx.nl) acc = fma(P.csr_val[e], S.x[(size_t)P.csr_col[e] * S.ldx], acc);
        acc = cx.warp_sum(acc);
        const double gi = K.g[i];
        double res = fabs(acc - gi) / (1.0 + fabs(gi));
        if (res != res) res = INFINITY;
        if (cx.lane == 0 && (bi < 0 || res > worst)) { worst = res; bi = i; }
    }
    cx.argmax(worst, bi, aux);
    return worst;
}

// Crossover to a vertex (north_star: "a batched restarted-PDHG path ... followed by crossover to a vertex").  x0 is a point
// that is optimal within the first-order method's tolerance (structural values of the scaled LP, element stride ldx0).
//  1. every variable is classified against its bounds: on a bound -> non-basic there, strictly between -> wanted basic;
//  2. the basis is crashed: starting from the scenario's vertex of the previous timestep (or from the slack basis after a
//     reset) every wanted structural that is not basic is pivoted in against a basic variable that should sit on a bound
//     (largest pivot element; a structural without an acceptable pivot stays non-basic at its nearer bound, a variable that
//     could not leave stays basic) - m basic variables, non-singular by construction, a handful of pivots per timestep;
//  3. the simplex of this file finishes from that basis (dual phase with shifted costs, primal clean-up): normally a handful
//     of pivots, because the crashed basis is the optimal one up to degenerate ties.
// The result is a vertex with a GLPK-style basis (b200lp_get_basis), like the one GLPK hands to BasisManager
// (cython_glpk.pyx:197-232).  Falls back to a cold start of the simplex if the crashed basis fails.
template <class CX>
RSX_FN int crossover(CX &cx, const Lp &P, const Scen &S, Work &K, const double *x0, const size_t ldx0, int &pivots, int &refactors,
                     int &crashed) {
    const int m = P.m, n = P.n, nv = n + m;
    load_bounds(cx, P, S, K);
    // row activities of x0 into g
    for (int i = cx.tid; i < m; i += cx.nt) {
        const int e0 = RSX_RO(&P.csr_ptr[i]), e1 = RSX_RO(&P.csr_ptr[i + 1]);
        if (e1 - e0 > P.long_len) continue;
        double acc = 0.0;
        for (int e = e0; e < e1; e++) acc = fma(RSX_RO(&P.csr_val[e]), x0[(size_t)RSX_RO(&P.csr_col[e]) * ldx0], acc);
        K.g[i] = acc;
    }
    for (int li = cx.warp; li < P.n_long; li += cx.nw) {
        const int i = P.long_rows[li];
        double acc = 0.0;
        for (int e = P.csr_ptr[i] + cx.lane; e < P.csr_ptr[i + 1]; e += cx.nl) acc = fma(P.csr_val[e], x0[(size_t)P.csr_col[e] * ldx0], acc);
        acc = cx.warp_sum(acc);
        if (cx.lane == 0) K.g[i] = acc;
    }
    cx.sync();
    // 1. classification: the status every variable should have (in dw, as a number) and its value at x0 (in alpha)
    for (int v = cx.tid; v < nv; v += cx.nt) {
        const double val = v < n ? x0[(size_t)v * ldx0] : K.g[v - n];
        const double lo = K.LO[v], hi = K.HI[v];
        const double tl = 1e-6 * (1.0 + fabs(lo)), th = 1e-6 * (1.0 + fabs(hi));
        int st;
        if (lo == hi) st = V_NS;
        else if (lo > -INFINITY && val <= lo + tl) st = V_NL;
        else if (hi < INFINITY && val >= hi - th) st = V_NU;
        else st = V_BASIC;
        K.dw[v] = (double)st;
        K.alpha[v] = val;
    }
    // 2. crash.  Warm: the scenario's vertex of the previous timestep (its W is the inverse of that basis) is repaired - only the
    // structurals that should be basic and are not get pivoted in, against basic variables that should sit on a bound.  Cold
    // (first timestep after a reset): the same from the slack basis, W = -I.
    const bool warm = S.info[0] != 0;
    if (warm) {
        for (int v = cx.tid; v < nv; v += cx.nt) K.stat[v] = S.stat[v];
        for (int i = cx.tid; i < m; i += cx.nt) K.head[i] = S.head[i];
    } else {
        for (int v = cx.tid; v < nv; v += cx.nt) K.stat[v] = (unsigned char)(v < n ? V_NL : V_BASIC);
        for (int i = cx.tid; i < m; i += cx.nt) K.head[i] = n + i;
        const size_t tot = (size_t)m * P.ldw;
        for (size_t e = cx.tid; e < tot; e += cx.nt) S.W[e] = 0.0;
        cx.sync();
        for (int i = cx.tid; i < m; i += cx.nt) S.W[(size_t)i * P.ldw + i] = -1.0;
    }
    cx.sync();
    int entered = 0;
    for (int q = 0; q < n; q++) {
        if ((int)K.dw[q] != V_BASIC || K.stat[q] == V_BASIC) continue;
        compute_w(cx, P, S, K, q);
        double best = 1e-7; int bi = -1, aux = 0;
        for (int i = cx.tid; i < m; i += cx.nt) {
            if ((int)K.dw[K.head[i]] == V_BASIC) continue;   // this basic variable is where it should be
            const double a = fabs(K.w[i]);
            if (a > best) { best = a; bi = i; }
        }
        if (bi < 0) best = -1.0;
        cx.argmax(best, bi, aux);
        if (bi < 0) continue;   // nothing can give way: the structural stays non-basic (status set below)
        load_rho(cx, P, S, K, bi);
        update_W(cx, P, S, K, bi);
        if (cx.leader()) {
            const int p = K.head[bi];
            K.stat[p] = (unsigned char)(int)K.dw[p];
            K.head[bi] = q;
            K.stat[q] = V_BASIC;
        }
        cx.sync();
        entered++;
    }
    // non-basic variables go where the point has them; one that should be basic but could not enter goes to its nearer bound
    for (int v = cx.tid; v < nv; v += cx.nt) {
        if (K.stat[v] == V_BASIC) continue;
        int st = (int)K.dw[v];
        if (st == V_BASIC) {
            const double lo = K.LO[v], hi = K.HI[v], val = K.alpha[v];
            st = !(lo > -INFINITY) ? (hi < INFINITY ? V_NU : V_NF) : (!(hi < INFINITY) ? V_NL : (val - lo <= hi - val ? V_NL : V_NU));
        }
        K.stat[v] = (unsigned char)st;
    }
    cx.sync();
    crashed += entered;
    // 3. finish with the simplex
    compute_d(cx, P, S, K);
    const int p_before = pivots;
    int status = solve(cx, P, S, K, pivots);
    if (status == R_OPTIMAL) {
        const double res = write_solution(cx, P, S, K);
        if (!(res <= 1e-7)) status = R_NUMERIC;
    }
    if (status != R_OPTIMAL) {  // cold start of the simplex (the retry ladder's last rung)
        refactors++;
        cold_init(cx, P, S, K);
        status = solve(cx, P, S, K, pivots);
        if (status == R_OPTIMAL) write_solution(cx, P, S, K);
    }
    for (int v = cx.tid; v < nv; v += cx.nt) { S.stat[v] = K.stat[v]; S.d[v] = K.d[v]; }
    for (int i = cx.tid; i < m; i += cx.nt) S.head[i] = K.head[i];
    if (cx.leader()) { S.info[0] = status == R_OPTIMAL ? 1 : 0; S.info[1] = pivots - p_before; }
    cx.sync();
    return status == R_NUMERIC ? R_SINGULAR : status;
}

// One scenario, one timestep: warm start from the stored basis (or the slack basis when `cold`), retry ladder
// (re-factorise the same basis, then restart from the slack basis), state written back.
template <class CX>
RSX_FN int run_scenario(CX &cx, const Lp &P, const Scen &S, Work &K, int cold, int &pivots, int &refactors) {
    const int m = P.m, nv = P.n + P.m;
    cx.mark(0);
    load_bounds(cx, P, S, K);
    cx.mark(1);
    const bool valid = !cold && S.info[0] != 0;
    if (valid) {
        for (int v = cx.tid; v < nv; v += cx.nt) { K.stat[v] = S.stat[v]; K.d[v] = S.d[v]; }
        for (int i = cx.tid; i < m; i += cx.nt) K.head[i] = S.head[i];
        cx.sync();
    }
    cx.mark(2);
    int status = R_OPTIMAL;
    const int p_before = pivots;
    const bool refresh = valid && P.refresh_pivots > 0 && S.info[1] >= P.refresh_pivots;
    for (int attempt = valid ? 0 : 2; attempt < 3; attempt++) {
        if (attempt == 1) {
            refactors++;
            if (refactor(cx, P, S, K) != 0) continue;
            compute_d(cx, P, S, K);
        } else if (attempt == 2) {
            cold_init(cx, P, S, K);
        } else if (P.cost_dynamic || refresh) {
            compute_d(cx, P, S, K);
        }
        status = solve(cx, P, S, K, pivots);
        cx.mark(6);
        if (status == R_OPTIMAL) {
            const double res = write_solution(cx, P, S, K);
            cx.mark(7);
            if (!(res <= 1e-7) && attempt < 2) status = R_NUMERIC;
        }
        if (status == R_OPTIMAL) break;
    }
    for (int v = cx.tid; v < nv; v += cx.nt) { S.stat[v] = K.stat[v]; S.d[v] = K.d[v]; }
    for (int i = cx.tid; i < m; i += cx.nt) S.head[i] = K.head[i];
    if (cx.leader()) {
        S.info[0] = status == R_OPTIMAL ? 1 : 0;
        S.info[1] = (refresh || !valid ? 0 : S.info[1]) + (pivots - p_before);   // pivots since the reduced costs were last rebuilt
    }
    cx.sync();
    cx.mark(8);
    return status == R_NUMERIC ? R_SINGULAR : status;
}

}  // namespace rsx
