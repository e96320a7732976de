// lp_kernel.cuh — sm_100a device code of the scenario-batched bounded simplex.
//
// One *group* of G lanes (G = 8, 16 or 32, a sub-warp) owns one scenario.  The dictionary T
// (m x n doubles, see below) lives in REGISTERS: lane l of the group holds rows l, l+G, ... (RPL rows
// per lane, NC padded columns, everything unrolled at compile time).  The small per-scenario vectors
// (values, bounds, non-basic values, reduced costs, the broadcast pivot row) live in shared memory;
// pricing and ratio tests are warp-shuffle arg-min / arg-max reductions over the group.  No tensor
// cores: there is no dense contraction anywhere on this path (BASELINE.json north_star).
//
// What it replaces in the reference: the body of CythonGLPKSolver._solve_scenario
// (pywr/solvers/cython_glpk.pyx:836-1095; edge form :1649-1918) for ALL scenarios of one timestep:
//   a3 objective update :880-897      a4 aggregated-factor coefficients :903-926
//   a5 row bounds :932-959            a6 storage / virtual-storage bounds :965-1019
//   a7 warm-started simplex from the scenario's own previous basis :197-232,1026-1063
//   a8 route/edge flow -> node flow :1068-1093
// and, in the device-resident run kernel, what Pywr does around the solver every timestep:
//   parameters (pywr/parameters/_parameters.pyx, _control_curves.pyx), Storage.after
//   (pywr/_core.pyx:1335-1361, a10) and the built-in recorders (pywr/recorders/_recorders.pyx, a11-a13).
// The algorithm is the one restated in oracle/lp_oracle.c (dense "dictionary" bounded dual simplex
// with cost shifting and a primal clean-up phase); every floating point operation that decides a
// pivot is performed in the same order with explicit fma(), and the file is compiled with
// --fmad=false, so that the two sides agree to the last bit on the LP solution.
//
// Dictionary: variables v = (x_0..x_{n-1}, r_0..r_{m-1}), r = A x.  m basic, n non-basic:
//     v_head[i] = sum_j T[i][j] * v_nb[j]
//
// Everything a timestep needs per scenario is addressed through ONE shared-memory value vector
//     V = [ slots (per-step inputs / op outputs) | constants | storage volumes | storage pc ]
// whose indices are resolved on the host (b200lp_create), so the per-timestep path has no
// "slot or constant?" branches and no global-memory indirections.
#pragma once
#ifndef __CUDACC_RTC__   // under NVRTC (jit_run.inc) the texts of b200lp.h and of this file are concatenated
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/b200lp.h"
#else
#ifndef INFINITY
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif
#ifndef NAN
#define NAN __longlong_as_double(0x7ff8000000000000LL)
#endif
#endif

#ifndef B200LP_SUBWARP_REDUX   // REDUX over a sub-warp member mask (groups of 4 / 8 / 16 lanes); 0: shuffle rounds for those
#define B200LP_SUBWARP_REDUX 1
#endif
#ifndef B200LP_RUN_MIN_BLOCKS
#define B200LP_RUN_MIN_BLOCKS 3
#endif

namespace b200lp {

#ifdef B200LP_PHASE_TIMING
// Per-phase cycle counters of scenario 0 (debug builds only: -DB200LP_PHASE_TIMING).  Accumulated in
// registers and written once at the end of the launch, so the probes cost a few cycles each.
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_INIT long long _pc = clock64(); unsigned long long _ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PHASE_MARK(k) do { const long long _c = clock64(); _ph[k] += (unsigned long long)(_c - _pc); _pc = _c; } while (0)
#define PHASE_FLUSH do { if (sidx == 0 && gl == 0) { for (int _k = 0; _k < 12; _k++) g_phase_cycles[_k] += _ph[_k]; } } while (0)
#else
#define PHASE_MARK(k)
#define PHASE_INIT
#define PHASE_FLUSH
#endif

constexpr double TOL_PRIMAL = 1e-9;
constexpr double TOL_DUAL = 1e-9;
constexpr double TOL_PIVOT = 1e-9;
constexpr double TIE_REL = 1e-9;
constexpr double TIE_ABS = 1e-12;
constexpr int REFACTOR_PIVOTS = 40;

enum : int { VS_BASIC = 1, VS_NL = 2, VS_NU = 3, VS_NF = 4, VS_NS = 5 };
enum : int { ROWC_STATIC = 0, ROWC_FLOW = 1, ROWC_STORAGE = 2, ROWC_PAD = 3 };
enum : int { K_SCRATCH = 0, K_ZERO = 1, K_ONE = 2, K_TWO = 3, K_PINF = 4, K_NINF = 5, K_COUNT = 6 };
// lane-parallel "fast ops" (one op per lane and dependency level)
enum : int { FOP_NONE = 0, FOP_PROD = 1, FOP_SUM = 2, FOP_MIN = 3, FOP_MAX = 4, FOP_CC = 5, FOP_INDEXED = 6, FOP_NEG = 7,
             FOP_MAX0 = 8, FOP_MIN0 = 9, FOP_DIV = 10 };
constexpr int FOP_MAXLEV = 4, FOP_WORDS = 8;  // kind, dst, a0..a5

// Shared LP structure, device pointers (built by b200lp_create).  All *_idx are indices into V.
struct DevStructure {
    int m, n, N, n_slots, n_consts, n_storage, n_dyn_a, cost_clip;
    int cost_dynamic;  // some cost term is a slot: reduced costs are rebuilt every step
    int rows_cap, nc;  // padded sizes of the kernel variant in use
    int n_val, vol_base, pc_base;
    int k_base;  // V[k_base + K_*]: scratch entry for idle lanes and built-in constants
    const double *A_dense;  // [nc][rows_cap] column-major, zero padded
    const int *row_class;   // [rows_cap] ROWC_*
    const int *row_a, *row_b;  // [rows_cap] V index of lb / ub, or (storage k, unused)
    const int *st_kind, *st_vol_idx, *st_minv_idx, *st_maxv_idx, *st_active_idx, *st_row;
    const int *cost_indptr, *cost_idx;
    const double *cost_w;
    const int *dyn_a_row, *dyn_a_col, *dyn_a_idx;
    const double *dyn_a_scale, *dyn_a_static;
    const int *flow_indptr, *flow_col;
    const double *flow_w;
    const double *consts;
};

// Per-scenario persistent state (the "basis" of BasisManager plus the cached dictionary).
struct DevState {
    double *T;    // [S][nc][rows_cap]
    double *d;    // [S][nc]
    int *head;    // [S][rows_cap]  variable in basic position i (-1: padding)
    int *nb;      // [S][nc]        variable in non-basic position j (-1: padding)
    int *nbstat;  // [S][nc]
    int *meta;    // [S][4]: valid, pivots since refactor, last status, unused
    double *vol;  // [n_storage][S] device-resident storage volume
    unsigned long long *counters;  // [0] pivots, [1] refactors
    int *fail;    // [0] number of failed scenarios, [1] min((scenario << 4) | code)
};


// minimum over the group of values that are >= 0 or +inf (ratios of the ratio tests).  A whole warp: the bit pattern of a
// non-negative double orders like the number, so two REDUX (high word, then low word among the lanes that hold the smallest
// high word) replace five dependent shuffle + DSETP rounds.
template <int G>
__device__ __forceinline__ double group_min(double v, unsigned mask) {
    if (G == 32 || B200LP_SUBWARP_REDUX) {   // sub-warp groups: every group reduces over its own member mask
        const unsigned gm = G == 32 ? 0xffffffffu : mask;
        const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);   // -0.0 -> +0.0
        const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
        const unsigned mh = __reduce_min_sync(gm, hi);
        const unsigned ml = __reduce_min_sync(gm, hi == mh ? lo : 0xffffffffu);
        return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, off, G));
    return v;
}

// arg-max of (key, lowest idx on ties); idx < 0 means "no candidate"
template <int G>
__device__ __forceinline__ void group_argmax(double &key, int &idx, int &aux, unsigned mask) {
    if (G == 32 || B200LP_SUBWARP_REDUX) {
        // keys of candidates are > 0 (bound violations, |pivot elements|, 1.0 under Bland's rule): bit patterns order like the
        // numbers.  Largest key by two REDUX, lowest index among the lanes that hold it by a third, winner's data by shuffles.
        const unsigned gm = G == 32 ? 0xffffffffu : mask;
        const bool cand = idx >= 0;
        const unsigned long long b = cand ? (unsigned long long)__double_as_longlong(key) + 1ull : 0ull;
        const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
        const unsigned mh = __reduce_max_sync(gm, hi);
        const unsigned ml = __reduce_max_sync(gm, hi == mh ? lo : 0u);
        const bool top = cand && hi == mh && lo == ml;
        const int best = (int)__reduce_min_sync(gm, top ? (unsigned)idx : 0xffffffffu);
        if (best < 0) { idx = -1; return; }   // no candidate anywhere (0xffffffff)
        const unsigned who = __ballot_sync(gm, top && idx == best) & gm;
        const int src = (__ffs(who) - 1) & (G - 1);   // lane within the group
        key = __shfl_sync(gm, key, src, G);
        aux = __shfl_sync(gm, aux, src, G);
        idx = best;
        return;
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        const double k2 = __shfl_xor_sync(mask, key, off, G);
        const int i2 = __shfl_xor_sync(mask, idx, off, G);
        const int a2 = __shfl_xor_sync(mask, aux, off, G);
        const bool take = (i2 >= 0) && (idx < 0 || k2 > key || (k2 == key && i2 < idx));
        if (take) { key = k2; idx = i2; aux = a2; }
    }
}

// OR of a small bit set over the group (one REDUX for a full warp)
template <int G>
__device__ __forceinline__ unsigned group_or(unsigned bits, unsigned mask) {
    if (G == 32) return __reduce_or_sync(0xffffffffu, bits);
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) bits |= __shfl_xor_sync(mask, bits, off, G);
    return bits;
}

// Tests on the bit pattern: integer compares have a quarter of the latency of DSETP on this part
// (benchmarks/tools/latency_probe.cu).  Results are identical to the floating point tests.
__device__ __forceinline__ bool is_nan_bits(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffull;
    return b > 0x7ff0000000000000ull;
}
__device__ __forceinline__ bool abs_below(double x, const unsigned long long thr_bits) {  // |x| < thr, false for NaN
    const unsigned long long b = (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffull;
    return b < thr_bits;
}
constexpr unsigned long long BITS_1E_8 = 0x3e45798ee2308c3aull;  // 1e-8

// a if lane == which else b, as ONE compare and one select (selp): written in PTX because the optimiser turns a chain of
// `if (lane == k) x = v_k;` into a tree of divergent branches
__device__ __forceinline__ double sel_lane(const int lane, const int which, const double a, const double b) {
    double r;
    asm("{ .reg .pred p; setp.eq.s32 p, %3, %4; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b), "r"(lane), "r"(which));
    return r;
}
__device__ __forceinline__ double sel_pred(const int c, const double a, const double b) {  // c != 0 ? a : b
    double r;
    asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b), "r"(c));
    return r;
}

// fmax(x, 0.0) as one compare + select: the same bits for every input (NaN -> 0, -0 -> +0, like CUDA's fmax)
__device__ __forceinline__ double max0(const double x) { return x > 0.0 ? x : 0.0; }

// cython_glpk.pyx:934-945 and constraint_type :66-77
__device__ __forceinline__ void type_bounds(double &lo, double &hi) {
    lo = abs_below(lo, BITS_1E_8) ? 0.0 : lo;
    hi = abs_below(hi, BITS_1E_8) ? 0.0 : hi;
    hi = abs_below(lo - hi, BITS_1E_8) ? lo : hi;
}

template <int G>
__device__ __forceinline__ unsigned group_ballot(int pred, unsigned mask) {
    if (G == 32) return __ballot_sync(0xffffffffu, pred);
    return __ballot_sync(mask, pred) & mask;
}

// Per-group shared memory: a fixed-layout block (compile-time offsets, so every access is one LDS/STS
// with an immediate offset) followed by the value vector V (run-time length).  Per-column arrays are
// padded to a multiple of G (NCP) so that "one column per lane" loops have no lane-dependent branch.
template <int G, int RPL, int NC>
struct GroupFixed {
    static constexpr int ROWS = G * RPL;
    static constexpr int NB = 1 + NC + ROWS;  // bounds arrays are indexed by variable id + 1
    static constexpr int NCP = ((NC + G - 1) / G) * G;
    double LO[NB], HI[NB];  // bounds of every variable (entry 0: padding variable, free)
    double xn[NCP];         // value of each non-basic position
    double d[NCP], dw[NCP]; // true / working reduced costs
    double prow[NCP];       // broadcast pivot row
    double x[NCP];          // column values by variable id
    double cost[NCP];
    double ract[ROWS];      // row activities by row id
    int nb[NCP], nbstat[NCP];
    int vstat[NCP + ROWS];  // scratch for the refactorisation
};

template <int G, int RPL, int NC>
struct GroupSmem {
    static constexpr int ROWS = G * RPL;
    static constexpr int NCP = GroupFixed<G, RPL, NC>::NCP;
    GroupFixed<G, RPL, NC> *f;
    double *V;  // [n_val + K_COUNT] slots | consts | vol | pc | scratch + built-in constants (K_*)
    __host__ __device__ static size_t bytes(int n_val) {
        return ((sizeof(GroupFixed<G, RPL, NC>) + 15) & ~(size_t)15) + sizeof(double) * (size_t)(n_val + K_COUNT + 1);
    }
    __device__ void carve(unsigned char *base, int n_val) {
        f = reinterpret_cast<GroupFixed<G, RPL, NC> *>(base);
        V = reinterpret_cast<double *>(base + ((sizeof(GroupFixed<G, RPL, NC>) + 15) & ~(size_t)15));
        (void)n_val;
    }
};

// ------------------------------------------------------------------------------------------------
// The per-scenario solver.  All members are per-lane registers except the GroupSmem pointers.
// ------------------------------------------------------------------------------------------------
template <int G, int RPL, int NC>
struct Dict {
    static constexpr int ROWS = G * RPL;
    static constexpr int NCP = GroupFixed<G, RPL, NC>::NCP;
    double T[RPL][NC];
    int head[RPL];
    double lob[RPL], hib[RPL];  // bounds of the basic variable in each owned position
    double beta[RPL];
    GroupSmem<G, RPL, NC> sm;
    unsigned mask;
    int gl;  // lane within the group
    int n, m;
    int pivots;

    __device__ __forceinline__ unsigned gmask() const { return G == 32 ? 0xffffffffu : mask; }
    __device__ __forceinline__ void sync() const {
        if (G == 32) __syncwarp(); else __syncwarp(mask);
    }

    __device__ __forceinline__ void load_basic_bounds() {
#pragma unroll
        for (int r = 0; r < RPL; r++) { lob[r] = sm.f->LO[head[r] + 1]; hib[r] = sm.f->HI[head[r] + 1]; }
    }

    __device__ __forceinline__ double col_of(int r, int q) const {
        double f = 0.0;
#pragma unroll
        for (int j = 0; j < NC; j++) f = (j == q) ? T[r][j] : f;
        return f;
    }

    // beta_i = sum_j T[i][j] xn[j], j ascending (same order as dict_basic_values in the oracle)
    __device__ __forceinline__ void basic_values() {
        double xv[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) xv[j] = sm.f->xn[j];
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < NC; j++) acc = fma(T[r][j], xv[j], acc);
            beta[r] = acc;
        }
    }

    // Exchange basic position (lane pl, slot ps) with non-basic position q.  prow must already hold
    // the pivot row in shared memory.  use_dw: also carry the working reduced costs.
    __device__ __forceinline__ void pivot(int pl, int ps, int q, bool use_dw) {
        const double inv = 1.0 / sm.f->prow[q];
        double pr[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) pr[j] = (j == q) ? inv : -sm.f->prow[j] * inv;
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            const bool is_piv = (gl == pl) && (r == ps);
            const double f = col_of(r, q);
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double tn = (j == q) ? f * inv : fma(f, pr[j], T[r][j]);
                T[r][j] = is_piv ? pr[j] : tn;
            }
        }
        // reduced-cost rows, one column per lane
        const double fd = sm.f->d[q], fw = sm.f->dw[q];
        const int vin = sm.f->nb[q];
        int vout = 0;
        {
            int hv = 0;
#pragma unroll
            for (int r = 0; r < RPL; r++) hv = (r == ps) ? head[r] : hv;
            vout = __shfl_sync(gmask(), hv, pl, G);
        }
        sync();
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
            const double prj = (j == q) ? inv : -sm.f->prow[j] * inv;
            sm.f->d[j] = (j == q) ? fd * inv : fma(fd, prj, sm.f->d[j]);
            if (use_dw) sm.f->dw[j] = (j == q) ? fw * inv : fma(fw, prj, sm.f->dw[j]);
        }
        if (gl == pl) {
#pragma unroll
            for (int r = 0; r < RPL; r++)
                if (r == ps) head[r] = vin;
        }
        if (gl == 0) sm.f->nb[q] = vout;
        pivots++;
        sync();
    }

    // lane pl publishes its row ps to shared memory
    __device__ __forceinline__ void publish_row(int pl, int ps) {
        if (gl == pl) {
#pragma unroll
            for (int r = 0; r < RPL; r++)
                if (r == ps) {
#pragma unroll
                    for (int j = 0; j < NC; j++) sm.f->prow[j] = T[r][j];
                }
        }
        sync();
    }

    // d_j = c_nb[j] + sum_i c_head[i] T[i][j]
    __device__ __forceinline__ void reduced_costs() {
        double part[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) part[j] = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            const int v = head[r];
            const double c = (v >= 0 && v < n) ? sm.f->cost[v] : 0.0;
#pragma unroll
            for (int j = 0; j < NC; j++) part[j] = fma(c, T[r][j], part[j]);
        }
#pragma unroll
        for (int j = 0; j < NC; j++) {
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) part[j] += __shfl_xor_sync(gmask(), part[j], off, G);
        }
        sync();
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
            const int v = sm.f->nb[j];
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < NC; k++) acc = (k == j) ? part[k] : acc;
            sm.f->d[j] = ((v >= 0 && v < n) ? sm.f->cost[v] : 0.0) + acc;
        }
        sync();
    }

    __device__ __forceinline__ void slack_basis() {
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) { sm.f->nb[j] = j < n ? j : -1; sm.f->nbstat[j] = j < n ? VS_NL : VS_NS; }
#pragma unroll
        for (int r = 0; r < RPL; r++) { const int i = r * G + gl; head[r] = i < m ? n + i : -1; }
        sync();
    }

    // Rebuild T for the current basis by Gauss-Jordan pivoting from the slack basis (dict_refactor
    // in the oracle).  Returns 0, or -1 if singular.
    __device__ __forceinline__ int refactor(const DevStructure &st) {
        // record the wanted status of every variable
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
            const int v = sm.f->nb[j];
            if (v >= 0) sm.f->vstat[v < n ? v : NCP + (v - n)] = sm.f->nbstat[j];
        }
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            const int v = head[r];
            if (v >= 0) sm.f->vstat[v < n ? v : NCP + (v - n)] = VS_BASIC;
        }
        sync();
        // slack dictionary T = A (with this scenario's dynamic coefficients)
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            const int i = r * G + gl;
#pragma unroll
            for (int j = 0; j < NC; j++) T[r][j] = __ldg(&st.A_dense[(size_t)j * ROWS + i]);
            head[r] = i < m ? n + i : -1;
        }
        for (int k = 0; k < st.n_dyn_a; k++) {
            const int row = __ldg(&st.dyn_a_row[k]), col = __ldg(&st.dyn_a_col[k]);
            const double v = __ldg(&st.dyn_a_scale[k]) * sm.V[__ldg(&st.dyn_a_idx[k])];
            const double delta = v - __ldg(&st.dyn_a_static[k]);
            const bool mine = (row % G) == gl;
#pragma unroll
            for (int r = 0; r < RPL; r++) {
#pragma unroll
                for (int j = 0; j < NC; j++)  // select, not an indexed store: T must stay in registers
                    T[r][j] = (mine && r == row / G && j == col) ? T[r][j] + delta : T[r][j];
            }
        }
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) { sm.f->nb[j] = j < n ? j : -1; sm.f->d[j] = 0.0; }
        sync();
        int ok = 0;
#pragma unroll 1
        for (int q = 0; q < n; q++) {
            if (sm.f->vstat[q] != VS_BASIC) continue;
            double best = TOL_PIVOT; int bi = -1, aux = 0;
#pragma unroll
            for (int r = 0; r < RPL; r++) {
                const int v = head[r];
                if (v < n) continue;  // padding (-1) or a structural already pivoted in
                if (sm.f->vstat[NCP + (v - n)] == VS_BASIC) continue;
                const double a = fabs(col_of(r, q));
                if (a > best) { best = a; bi = r * G + gl; }
            }
            if (bi < 0) best = -1.0;
            group_argmax<G>(best, bi, aux, gmask());
            if (bi < 0) { ok = -1; break; }
            publish_row(bi % G, bi / G);
            const int save = pivots;
            pivot(bi % G, bi / G, q, false);
            pivots = save;
        }
        // statuses of the non-basic positions
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
            const int v = sm.f->nb[j];
            sm.f->nbstat[j] = v < 0 ? VS_NS : sm.f->vstat[v < n ? v : NCP + (v - n)];
        }
        sync();
        return ok;
    }

    // status + value of every non-basic position for the current bounds (dict_place_nonbasics),
    // written with selects: all lanes follow one instruction stream
    __device__ __forceinline__ int place_nonbasics_lane() {
        int infeas = 0;
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
            const int v = sm.f->nb[j];
            const double lo = sm.f->LO[v + 1], hi = sm.f->HI[v + 1], dj = sm.f->d[j];
            const int s0 = sm.f->nbstat[j];
            const bool lo_fin = lo > -INFINITY, hi_fin = hi < INFINITY, fx = lo == hi;
            const bool dpos = dj > TOL_DUAL, dneg = dj < -TOL_DUAL;
            const int s_box = dpos ? VS_NL : (dneg ? VS_NU : ((s0 == VS_NL || s0 == VS_NU) ? s0 : VS_NL));
            int s = lo_fin ? (hi_fin ? s_box : VS_NL) : (hi_fin ? VS_NU : VS_NF);
            s = fx ? VS_NS : s;
            const bool bad = lo_fin ? (!hi_fin && dneg) : (hi_fin ? dpos : (dpos || dneg));
            double xv = (s == VS_NU) ? hi : (s == VS_NF ? 0.0 : lo);
            if (v < 0) { s = VS_NS; xv = 0.0; }
            infeas |= (!fx && bad && v >= 0) ? 1 : 0;
            sm.f->nbstat[j] = s;
            sm.f->xn[j] = xv;
        }
        return infeas;
    }

    __device__ __forceinline__ int place_nonbasics() {
        int infeas = place_nonbasics_lane();
        infeas = group_ballot<G>(infeas, gmask()) != 0u;
        sync();
        return infeas;
    }

    // Steady-state timestep: place the non-basics, evaluate the basics and decide with ONE group
    // reduction whether anything needs the simplex.  `lane_flags`: bit 0 NaN input, bit 1 inconsistent
    // bounds (from assemble_rows).  Returns a status, or -1 when the general solve() must run.
    __device__ __forceinline__ int fast_step(const int lane_flags) {
        const int infe = place_nonbasics_lane();
        sync();
        load_basic_bounds();
        basic_values();
        int viol = 0;
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            const double b = beta[r];
            viol |= (b < lob[r] - TOL_PRIMAL * (1.0 + fabs(lob[r]))) || (b > hib[r] + TOL_PRIMAL * (1.0 + fabs(hib[r])));
        }
        const unsigned all = group_or<G>((unsigned)lane_flags | (infe << 2) | (viol << 3), gmask());
        if (all == 0u) return B200LP_ST_OPTIMAL;
        if (all & 1u) return B200LP_ST_NAN;
        if (all & 2u) return B200LP_ST_INFEASIBLE;
        return -1;
    }

    // working reduced costs of the (possibly cost-shifted) dual phase
    __device__ __forceinline__ void init_dw(int shifted) {
        _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
            double w = sm.f->d[j];
            if (shifted) {
                const int s = sm.f->nbstat[j];
                if ((s == VS_NL && w < -TOL_DUAL) || (s == VS_NU && w > TOL_DUAL) || (s == VS_NF && fabs(w) > TOL_DUAL))
                    w = 0.0;
            }
            sm.f->dw[j] = w;
        }
        sync();
    }

    // dict_solve of the oracle.  Returns a B200LP_ST_* code (uniform over the group).
    __device__ __forceinline__ int solve() {
        const int cap_bland = 20 * (m + n) + 50, cap_fail = 200 * (m + n) + 1000;
        int shifted = place_nonbasics();
        bool dw_ready = false;
        load_basic_bounds();
        int phase = 0;
#pragma unroll 1
        for (int it = 0;; it++) {
            if (it > cap_fail) return B200LP_ST_ITERLIMIT;
            const bool bland = it > cap_bland;
            basic_values();
            int pl, ps, q, to_upper;
            if (phase == 0) {
                // ---- leaving variable: largest bound violation -------------------------------
                double worst = -1.0; int bi = -1, dir = 0;
#pragma unroll
                for (int r = 0; r < RPL; r++) {
                    const double b = beta[r];
                    double viol = 0.0; int dd = 0;
                    if (b < lob[r] - TOL_PRIMAL * (1.0 + fabs(lob[r]))) { viol = lob[r] - b; dd = +1; }
                    else if (b > hib[r] + TOL_PRIMAL * (1.0 + fabs(hib[r]))) { viol = b - hib[r]; dd = -1; }
                    if (dd != 0) {
                        if (bland) viol = 1.0;
                        if (bi < 0 || viol > worst) { worst = viol; bi = r * G + gl; dir = dd; }
                    }
                }
                if (group_ballot<G>(bi >= 0, gmask()) == 0u) {  // primal feasible
                    if (!shifted) return B200LP_ST_OPTIMAL;
                    phase = 1; shifted = 0;
                    continue;
                }
                group_argmax<G>(worst, bi, dir, gmask());
                if (!dw_ready) { init_dw(shifted); dw_ready = true; }
                pl = bi % G; ps = bi / G;
                publish_row(pl, ps);
                // ---- dual ratio test ------------------------------------------------------------
                double rmin = INFINITY;
                double ratio1 = INFINITY, a1 = 0.0;   // NCP == G: this lane's one column, kept for the second pass
                _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
                    const int s = sm.f->nbstat[j];
                    const double a = sm.f->prow[j] * (double)dir;
                    bool ok = false;
                    if (s == VS_NL) ok = a > TOL_PIVOT;
                    else if (s == VS_NU) ok = a < -TOL_PIVOT;
                    else if (s == VS_NF) ok = fabs(a) > TOL_PIVOT;
                    if (ok) { ratio1 = fabs(sm.f->dw[j]) / fabs(a); a1 = a; rmin = fmin(rmin, ratio1); }
                }
                rmin = group_min<G>(rmin, gmask());
                if (rmin == INFINITY) return B200LP_ST_INFEASIBLE;
                const double thr = rmin * (1.0 + TIE_REL) + TIE_ABS;
                double amax = -1.0; int aux = 0;
                q = -1;
                if (NCP == G) {   // the same candidates and ratios as the first pass (one column per lane)
                    if (ratio1 < INFINITY && !(ratio1 > thr)) { amax = bland ? 1.0 : fabs(a1); q = gl; }
                } else {
                _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
                    const int s = sm.f->nbstat[j];
                    const double a = sm.f->prow[j] * (double)dir;
                    bool ok = false;
                    if (s == VS_NL) ok = a > TOL_PIVOT;
                    else if (s == VS_NU) ok = a < -TOL_PIVOT;
                    else if (s == VS_NF) ok = fabs(a) > TOL_PIVOT;
                    if (!ok) continue;
                    if (fabs(sm.f->dw[j]) / fabs(a) > thr) continue;
                    const double key = bland ? 1.0 : fabs(a);
                    if (q < 0 || key > amax) { amax = key; q = j; }
                }
                }
                group_argmax<G>(amax, q, aux, gmask());
                to_upper = dir < 0;  // leaving variable goes to the bound it violated
            } else {
                // ---- entering variable: largest dual infeasibility -----------------------------
                double worst = -1.0; int dir = 0;
                q = -1;
                _Pragma("unroll") for (int j = gl; j < NCP; j += G) {
                    const int s = sm.f->nbstat[j];
                    const double dj = sm.f->d[j];
                    int dd = 0;
                    if ((s == VS_NL || s == VS_NF) && dj < -TOL_DUAL) dd = +1;
                    else if ((s == VS_NU || s == VS_NF) && dj > TOL_DUAL) dd = -1;
                    if (dd != 0) {
                        const double key = bland ? 1.0 : fabs(dj);
                        if (q < 0 || key > worst) { worst = key; q = j; dir = dd; }
                    }
                }
                group_argmax<G>(worst, q, dir, gmask());
                if (q < 0) return B200LP_ST_OPTIMAL;
                const int vin = sm.f->nb[q];
                const double inlo = sm.f->LO[vin + 1], inhi = sm.f->HI[vin + 1];
                double tmin = inhi - inlo;
                if (!(tmin < INFINITY)) tmin = INFINITY;
                // ---- primal ratio test down column q ---------------------------------------------
                double step_min = INFINITY;
                double stp[RPL], aa[RPL]; int upf[RPL];
#pragma unroll
                for (int r = 0; r < RPL; r++) {
                    const double a = col_of(r, q) * (double)dir;
                    double step = INFINITY; int up = 0;
                    if (a > TOL_PIVOT) { if (hib[r] < INFINITY) { step = (hib[r] - beta[r]) / a; up = 1; } }
                    else if (a < -TOL_PIVOT) { if (lob[r] > -INFINITY) step = (lob[r] - beta[r]) / a; }
                    if (step < 0.0) step = 0.0;
                    stp[r] = step; aa[r] = fabs(a); upf[r] = up;
                    step_min = fmin(step_min, step);
                }
                step_min = group_min<G>(step_min, gmask());
                if (step_min == INFINITY && tmin == INFINITY) return B200LP_ST_UNBOUNDED;
                if (tmin <= step_min) {
                    if (gl == 0) {
                        sm.f->nbstat[q] = dir > 0 ? VS_NU : VS_NL;
                        sm.f->xn[q] = dir > 0 ? inhi : inlo;
                    }
                    sync();
                    continue;
                }
                const double thr = step_min * (1.0 + TIE_REL) + TIE_ABS;
                double amax = -1.0; int bi = -1, up = 0;
#pragma unroll
                for (int r = 0; r < RPL; r++) {
                    if (stp[r] > thr) continue;
                    const double key = bland ? 1.0 : aa[r];
                    if (bi < 0 || key > amax) { amax = key; bi = r * G + gl; up = upf[r]; }
                }
                group_argmax<G>(amax, bi, up, gmask());
                pl = bi % G; ps = bi / G;
                publish_row(pl, ps);
                to_upper = up;
            }
            // ---- the pivot (single call site for both phases) ---------------------------------
            double vlo = 0.0, vhi = 0.0;
#pragma unroll
            for (int r = 0; r < RPL; r++) { vlo = (r == ps) ? lob[r] : vlo; vhi = (r == ps) ? hib[r] : vhi; }
            vlo = __shfl_sync(gmask(), vlo, pl, G);
            vhi = __shfl_sync(gmask(), vhi, pl, G);
            pivot(pl, ps, q, phase == 0);
            if (gl == 0) {
                sm.f->nbstat[q] = (vlo == vhi) ? VS_NS : (to_upper ? VS_NU : VS_NL);
                sm.f->xn[q] = to_upper ? vhi : vlo;
            }
            sync();
            load_basic_bounds();
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Per-timestep pipeline pieces shared by the step kernel and the run kernel.
// ------------------------------------------------------------------------------------------------

template <int G, int RPL, int NC>
__device__ __forceinline__ void init_group(Dict<G, RPL, NC> &D, const DevStructure &st, unsigned char *smem_raw,
                                           const int gib) {
    D.gl = threadIdx.x % G;
    const int lane = threadIdx.x & 31;
    D.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    D.n = st.n; D.m = st.m; D.pivots = 0;
    const size_t gbytes = (GroupSmem<G, RPL, NC>::bytes(st.n_val) + 15) & ~(size_t)15;
    D.sm.carve(smem_raw + gbytes * gib, st.n_val);
    for (int j = NC + D.gl; j < D.NCP; j += G) {  // padding columns: never eligible, all zeros
        D.sm.f->nb[j] = -1; D.sm.f->nbstat[j] = VS_NS;
        D.sm.f->d[j] = 0.0; D.sm.f->dw[j] = 0.0; D.sm.f->prow[j] = 0.0; D.sm.f->xn[j] = 0.0;
        D.sm.f->x[j] = 0.0; D.sm.f->cost[j] = 0.0;
    }
}

// Row descriptors of the rows a lane owns (registers, loaded once per launch).  Every lane evaluates
// BOTH bound formulas each timestep and selects (no divergent branches): a/b are the V indices of
// lb/ub for flow rows (also for rows whose bounds are constants), vol/minv/maxv/act those of the
// storage formula (pointing at harmless entries for the other rows).
template <int RPL>
struct RowDesc {
    int cls[RPL], a[RPL], b[RPL], vol[RPL], minv[RPL], maxv[RPL], act[RPL];
};

// constants -> V, bounds of columns / padding.  Called once per launch.
template <int G, int RPL, int NC>
__device__ __forceinline__ void prepare_static(const DevStructure &st, Dict<G, RPL, NC> &D, RowDesc<RPL> &rd) {
    const int gl = D.gl, n = st.n;
    for (int k = gl; k < st.n_consts; k += G) D.sm.V[st.n_slots + k] = __ldg(&st.consts[k]);
    if (gl == 0) {
        double *K = D.sm.V + st.k_base;
        K[K_SCRATCH] = 0.0; K[K_ZERO] = 0.0; K[K_ONE] = 1.0; K[K_TWO] = 2.0; K[K_PINF] = INFINITY; K[K_NINF] = -INFINITY;
    }
    for (int v = gl; v < 1 + NC; v += G) {  // entry 0 (padding variable) and the columns x >= 0
        const bool col = v >= 1 && v <= n;
        D.sm.f->LO[v] = col ? 0.0 : -INFINITY;
        D.sm.f->HI[v] = INFINITY;
    }
#pragma unroll
    for (int r = 0; r < RPL; r++) {
        const int i = r * G + gl;
        rd.cls[r] = __ldg(&st.row_class[i]);
        rd.a[r] = __ldg(&st.row_a[i]);
        rd.b[r] = __ldg(&st.row_b[i]);
        rd.vol[r] = rd.minv[r] = rd.maxv[r] = rd.act[r] = 0;
        if (rd.cls[r] == ROWC_STORAGE) {
            const int k = rd.a[r];
            rd.vol[r] = __ldg(&st.st_vol_idx[k]); rd.minv[r] = __ldg(&st.st_minv_idx[k]);
            rd.maxv[r] = __ldg(&st.st_maxv_idx[k]); rd.act[r] = __ldg(&st.st_active_idx[k]);
            rd.a[r] = rd.b[r] = 0;
        }
        if (rd.cls[r] == ROWC_PAD) {  // padding rows are free
            D.sm.f->LO[1 + n + i] = -INFINITY; D.sm.f->HI[1 + n + i] = INFINITY;
        }
    }
    D.sync();
}

// (a3) objective c_j = sum w * V[idx], clipped (cython_glpk.pyx:886-897)
template <int G, int RPL, int NC>
__device__ __forceinline__ int assemble_costs(const DevStructure &st, Dict<G, RPL, NC> &D) {
    int nan_seen = 0;
    for (int j = D.gl; j < NC; j += G) {
        double acc = 0.0;
        if (j < st.n) {
            for (int k = __ldg(&st.cost_indptr[j]); k < __ldg(&st.cost_indptr[j + 1]); k++)
                acc += __ldg(&st.cost_w[k]) * D.sm.V[__ldg(&st.cost_idx[k])];
            if (isnan(acc)) nan_seen = 1;
            if (st.cost_clip && fabs(acc) < 1e-8) acc = 0.0;
        }
        D.sm.f->cost[j] = acc;
    }
    return nan_seen;
}

// (a5/a6) bounds of every row from the value vector V.  Branch-free per lane: both the flow formula
// (cython_glpk.pyx:932-959) and the storage formula (:965-1019) are evaluated and the row's class
// selects.  Returns this LANE's flags (bit 0 NaN input, bit 1 inconsistent bounds); the caller reduces
// them over the group together with the feasibility test (Dict::fast_step).
template <int G, int RPL, int NC>
__device__ __forceinline__ int assemble_rows(const DevStructure &st, Dict<G, RPL, NC> &D, const RowDesc<RPL> &rd,
                                             const double days, int flags) {
    const int n = st.n;
    const bool unit_days = days == 1.0;  // x / 1.0 == x exactly: skip the divisions on daily timesteps
#pragma unroll
    for (int r = 0; r < RPL; r++) {
        const int i = r * G + D.gl;
        // flow formula
        const double fl = D.sm.V[rd.a[r]], fu = D.sm.V[rd.b[r]];
        const bool fnan = is_nan_bits(fl) || is_nan_bits(fu);
        // storage formula
        const double maxv = D.sm.V[rd.maxv[r]], minv = D.sm.V[rd.minv[r]], v = D.sm.V[rd.vol[r]];
        const bool active = D.sm.V[rd.act[r]] != 0.0;  // plain storages point at a constant 1
        const bool snan = is_nan_bits(maxv) || is_nan_bits(minv) || is_nan_bits(v);
        const double avail = fmax(v - minv, 0.0), room = fmax(maxv - v, 0.0);
        double sl = -avail, su = room;
        if (!unit_days) { sl = -avail / days; su = room / days; }
        const bool is_st = rd.cls[r] == ROWC_STORAGE;
        double l = is_st ? sl : fl, u = is_st ? su : fu;
        type_bounds(l, u);  // clipping / FX typing is the same for both (cython_glpk.pyx:934-945,981-984)
        const bool freed = is_st && !active, fixed0 = is_st && maxv == minv;
        l = freed ? -INFINITY : l; u = freed ? INFINITY : u;
        l = fixed0 ? 0.0 : l; u = fixed0 ? 0.0 : u;
        if (rd.cls[r] != ROWC_PAD) {
            if (is_st ? snan : fnan) flags |= 1;
            if (l > u) flags |= 2;
            D.sm.f->LO[1 + n + i] = l; D.sm.f->HI[1 + n + i] = u;
        }
    }
    if (st.cost_dynamic) flags |= assemble_costs(st, D);
    if (st.n_dyn_a > 0)
        for (int k = D.gl; k < st.n_dyn_a; k += G)
            if (isnan(D.sm.V[__ldg(&st.dyn_a_idx[k])])) flags |= 1;
    D.sync();
    return flags;
}

// (a7) the general solve with the oracle's retry ladder (role of retry_solve, cython_glpk.pyx:1034-1042): re-factorise when
// asked, run the simplex from the scenario's basis, on failure re-factorise the same basis, then restart from the slack basis.
template <int G, int RPL, int NC>
__device__ __forceinline__ int solve_lp_slow(const DevStructure &st, Dict<G, RPL, NC> &D, int &valid, int &npiv,
                                             int &refactors, bool need_refactor) {
    const bool fresh0 = need_refactor;
    int status = B200LP_ST_OPTIMAL;
#pragma unroll 1
    for (int attempt = 0; attempt < 3; attempt++) {
        if (attempt == 1 && fresh0) continue;
        if (attempt == 2) D.slack_basis();
        if (attempt > 0) need_refactor = true;
        if (need_refactor) {
            if (D.refactor(st) != 0) {
                D.slack_basis();
                D.refactor(st);
            }
            D.reduced_costs();
            npiv = 0; refactors++; valid = 1;
        } else if (st.cost_dynamic) {
            D.reduced_costs();
        }
        const int p0 = D.pivots;
        status = D.solve();
        npiv += D.pivots - p0;
        if (status == B200LP_ST_OPTIMAL) break;
    }
    if (status != B200LP_ST_OPTIMAL) valid = 0;
    return status;
}

// (a7) warm-started solve: the steady-state test first, the general solve when it does not settle the timestep.
// `lane_flags` are the per-lane input flags of assemble_rows.
template <int G, int RPL, int NC>
__device__ __forceinline__ int solve_lp(const DevStructure &st, Dict<G, RPL, NC> &D, int &valid, int &npiv,
                                        int &refactors, const int lane_flags) {
    const bool need_refactor = !valid || st.n_dyn_a > 0 || npiv >= REFACTOR_PIVOTS;
    if (!need_refactor && !st.cost_dynamic) {
        const int f = D.fast_step(lane_flags);  // 93 % of the benchmark's timesteps end here
        if (f >= 0) return f;
    } else {
        const unsigned all = group_or<G>((unsigned)lane_flags, D.gmask());
        if (all & 1u) return B200LP_ST_NAN;
        if (all & 2u) return B200LP_ST_INFEASIBLE;
    }
    return solve_lp_slow(st, D, valid, npiv, refactors, need_refactor);
}

// (a8) primal values of the columns into sm.f->x and the activities of the rows into sm.f->ract
template <int G, int RPL, int NC>
__device__ __forceinline__ void extract_solution(Dict<G, RPL, NC> &D) {
    // D.beta is current: solve() returns B200LP_ST_OPTIMAL right after basic_values() with the final
    // non-basic values (the oracle recomputes it from the same inputs: identical bits)
    const int n = D.n;
    _Pragma("unroll") for (int j = D.gl; j < D.NCP; j += G) {
        const int v = D.sm.f->nb[j];
        if (v >= n) D.sm.f->ract[v - n] = D.sm.f->xn[j];
        else if (v >= 0) D.sm.f->x[v] = D.sm.f->xn[j];
    }
#pragma unroll
    for (int r = 0; r < RPL; r++) {
        const int v = D.head[r];
        if (v >= n) D.sm.f->ract[v - n] = D.beta[r];
        else if (v >= 0) D.sm.f->x[v] = D.beta[r];
    }
    D.sync();
}

template <int G, int RPL, int NC>
__device__ __forceinline__ void load_state(const DevState &state, Dict<G, RPL, NC> &D, const int sidx, const int valid) {
    constexpr int ROWS = G * RPL;
    const double *Tg = state.T + (size_t)sidx * NC * ROWS;
#pragma unroll
    for (int r = 0; r < RPL; r++) {
        const int i = r * G + D.gl;
        D.head[r] = state.head[(size_t)sidx * ROWS + i];
        if (valid) {
#pragma unroll
            for (int j = 0; j < NC; j++) D.T[r][j] = Tg[(size_t)j * ROWS + i];
        }
    }
    for (int j = D.gl; j < NC; j += G) {
        D.sm.f->nb[j] = state.nb[(size_t)sidx * NC + j];
        D.sm.f->nbstat[j] = state.nbstat[(size_t)sidx * NC + j];
        D.sm.f->d[j] = state.d[(size_t)sidx * NC + j];
    }
}

template <int G, int RPL, int NC>
__device__ __forceinline__ void store_state(const DevState &state, Dict<G, RPL, NC> &D, const int sidx, const bool store_T) {
    constexpr int ROWS = G * RPL;
    double *Tg = state.T + (size_t)sidx * NC * ROWS;
    if (store_T) {
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            const int i = r * G + D.gl;
            state.head[(size_t)sidx * ROWS + i] = D.head[r];
#pragma unroll
            for (int j = 0; j < NC; j++) Tg[(size_t)j * ROWS + i] = D.T[r][j];
        }
    }
    for (int j = D.gl; j < NC; j += G) {
        state.nb[(size_t)sidx * NC + j] = D.sm.f->nb[j];
        state.nbstat[(size_t)sidx * NC + j] = D.sm.f->nbstat[j];
        state.d[(size_t)sidx * NC + j] = D.sm.f->d[j];
    }
}

// ------------------------------------------------------------------------------------------------
// One timestep for all scenarios (the per-timestep Solver.solve path).
// ------------------------------------------------------------------------------------------------
template <int G, int RPL, int NC>
__global__ void __launch_bounds__(128, (RPL == 1 ? 3 : 1))
lp_step_kernel(const DevStructure st, const DevState state, const int S, const double days,
               const double *__restrict__ slots, double *__restrict__ node_flow, double *__restrict__ col_flow,
               double *__restrict__ objective) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups_per_block = blockDim.x / G;
    const int gib = threadIdx.x / G;
    const int sidx = blockIdx.x * groups_per_block + gib;
    if (sidx >= S) return;  // whole group exits together

    Dict<G, RPL, NC> D;
    init_group(D, st, smem_raw, gib);
    const int gl = D.gl, n = st.n;

    for (int k = gl; k < st.n_slots; k += G) D.sm.V[k] = slots[(size_t)k * S + sidx];
    for (int k = gl; k < st.n_storage; k += G) D.sm.V[st.vol_base + k] = state.vol[(size_t)k * S + sidx];
    D.sync();
    RowDesc<RPL> rd;
    prepare_static(st, D, rd);
    int flags = 0;
    if (!st.cost_dynamic) flags |= assemble_costs(st, D);
    const int lane_flags = assemble_rows(st, D, rd, days, flags);

    int *meta = state.meta + (size_t)sidx * 4;
    int refactors = 0;
    int valid = meta[0], npiv = meta[1];
    load_state(state, D, sidx, valid);
    D.sync();
    int status = solve_lp(st, D, valid, npiv, refactors, lane_flags);
    // NaN / inconsistent inputs leave the scenario's basis untouched (the oracle returns before the solve)
    const bool input_error = status == B200LP_ST_NAN || (status == B200LP_ST_INFEASIBLE && D.pivots == 0 && refactors == 0 && valid);
    if (!input_error) {
        store_state(state, D, sidx, D.pivots > 0 || refactors > 0);
        if (gl == 0) { meta[0] = valid; meta[1] = npiv; }
    }
    if (gl == 0) {
        meta[2] = status;
        if (D.pivots) atomicAdd(&state.counters[0], (unsigned long long)D.pivots);
        if (refactors) atomicAdd(&state.counters[1], (unsigned long long)refactors);
        if (status != B200LP_ST_OPTIMAL) {
            atomicAdd(&state.fail[0], 1);
            atomicMin(&state.fail[1], (sidx << 4) | status);
        }
    }
    if (status != B200LP_ST_OPTIMAL) return;

    extract_solution(D);
    if (col_flow)
        for (int j = gl; j < n; j += G) col_flow[(size_t)j * S + sidx] = D.sm.f->x[j];
    if (objective && gl == 0) {
        double acc = 0.0;
        for (int j = 0; j < n; j++) acc = fma(D.sm.f->cost[j], D.sm.f->x[j], acc);
        objective[sidx] = acc;
    }
    if (node_flow) {
        for (int v = gl; v < st.N; v += G) {
            double acc = 0.0;
            for (int k = __ldg(&st.flow_indptr[v]); k < __ldg(&st.flow_indptr[v + 1]); k++)
                acc = fma(__ldg(&st.flow_w[k]), D.sm.f->x[__ldg(&st.flow_col[k])], acc);
            node_flow[(size_t)v * S + sidx] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Device-resident run: timesteps [t0, t0+nsteps) for all scenarios in ONE launch.  The dictionary
// stays in registers across timesteps; parameters, storage mass balance and recorders are evaluated
// here, so nothing returns to the host until read-out.
// ------------------------------------------------------------------------------------------------
struct DevProgram {
    int n_steps, n_tables, n_axes, n_recorders, n_table_ops, n_code;
    const double *ts_days;
    // table ops, one descriptor each (prefetched one timestep ahead by the lane that owns it)
    const int *tab_dst;      // [n_table_ops] V index written
    const int *tab_table;    // [n_table_ops]
    const int *tab_offset;   // [n_table_ops]
    const int *table_rows, *table_cols, *table_axis;
    const long long *table_offset;
    const double *table_data;
    const int *scen_index;
    // all other ops, pre-decoded: kind, dst, nargs, args... (V indices / storage ids / small ints)
    const int *code;
    // lane-parallel schedule of the same ops (n_fop_levels == 0: use the interpreter)
    int n_fop_levels;
    const int *fop;  // [n_fop_levels][32][FOP_WORDS]
    const int *stflow_indptr, *stflow_col;
    const double *stflow_w;
    const int *rec_kind, *rec_src, *rec_idx, *rec_row, *rec_indptr, *rec_col;
    const double *rec_factor, *rec_w;
    double *const *rec_out;  // [n_recorders] device outputs
    double *pc;              // [n_storage][S]
    const int *roll_len, *roll_off;  // [n_storage] ring-buffer length / first row of the rolling virtual storages
    double *roll_mem;        // [sum roll_len][S]
    double *last_x;          // [n][S]
    int *fail_step;          // [S]
    // temporal aggregates of the array recorders, accumulated in timestep order while the arrays are written:
    // [n_recorders][3][S] = sum, min, max over the timesteps run so far (Recorder.values() =
    // Aggregator.aggregate_2d(_data, axis=0), pywr/recorders/_recorders.pyx:109-184,630-633, without shipping _data)
    double *rec_agg;
};

// one sample of an array recorder into its running aggregates (NaN propagates like numpy's sum / min / max)
__device__ __forceinline__ void agg_sample(double &sum, double &mn, double &mx, const double o) {
    sum += o;
    mn = (o < mn || o != o) ? o : mn;
    mx = (o > mx || o != o) ? o : mx;
}

__device__ __forceinline__ double clamp_pc(double pc) {  // Storage.get_current_pc, pywr/_core.pyx:1027-1039
    if (pc > 1.0 || isnan(pc)) return 1.0;
    if (pc < 0.0) return 0.0;
    return pc;
}

// scalar interpreter for the non-table ops (lane 0 of the group); same arithmetic as orc_eval_ops.
// `code` is in shared memory; every operand is an index into V.
// VA: anything indexable that yields a double lvalue (a pointer into shared memory here, a strided plane view in
// plane_program.inc).
// where B200LP_OP_LAG finds yesterday's sample: the recorders' [T, S] output arrays, this scenario, this timestep
struct LagCtx {
    double *const *rec_out;
    int t, S, sidx;
};

template <class VA>
__device__ __forceinline__ void eval_ops(const int *code, const int n_code, VA V, const int pc_base, const LagCtx lag) {
    int pos = 0;
    while (pos < n_code) {
        const int kind = code[pos], dst = code[pos + 1], na = code[pos + 2];
        const int *a = code + pos + 3;
        pos += 3 + na;
        double v = 0.0;
        switch (kind) {
        case B200LP_OP_AGG: {
            const int func = a[0], cnt = a[1];
            if (func == B200LP_AGG_PRODUCT) { v = 1.0; for (int i = 0; i < cnt; i++) v *= V[a[2 + i]]; }
            else if (func == B200LP_AGG_SUM) { v = 0.0; for (int i = 0; i < cnt; i++) v += V[a[2 + i]]; }
            else if (func == B200LP_AGG_MAX) { v = -INFINITY; for (int i = 0; i < cnt; i++) { const double w = V[a[2 + i]]; if (w > v) v = w; } }
            else if (func == B200LP_AGG_MIN) { v = INFINITY; for (int i = 0; i < cnt; i++) { const double w = V[a[2 + i]]; if (w < v) v = w; } }
            else { v = 0.0; for (int i = 0; i < cnt; i++) v += V[a[2 + i]]; v /= cnt; }
        } break;
        case B200LP_OP_CONTROL_CURVE: {
            const int cnt = a[1];
            const double c = clamp_pc(V[pc_base + a[0]]);
            int idx = cnt;
            for (int i = 0; i < cnt; i++) if (c >= V[a[2 + i]]) { idx = i; break; }
            v = V[a[2 + cnt + idx]];
        } break;
        case B200LP_OP_CC_INDEX: {
            const int cnt = a[1];
            const double c = clamp_pc(V[pc_base + a[0]]);
            int idx = cnt;
            for (int i = 0; i < cnt; i++) if (c >= V[a[2 + i]]) { idx = i; break; }
            v = (double)idx;
        } break;
        case B200LP_OP_INDEXED: {
            int idx = (int)V[a[0]];
            if (idx < 0) idx = 0;
            if (idx > a[1] - 1) idx = a[1] - 1;
            v = V[a[2 + idx]];
        } break;
        case B200LP_OP_NEG: v = -V[a[0]]; break;
        case B200LP_OP_MAX0: { const double x = V[a[0]], th = V[a[1]]; v = x > th ? x : th; } break;
        case B200LP_OP_MIN0: { const double x = V[a[0]], th = V[a[1]]; v = x < th ? x : th; } break;
        case B200LP_OP_DIV: v = V[a[0]] / V[a[1]]; break;
        case B200LP_OP_STATE: v = a[1] ? clamp_pc(V[a[0]]) : V[a[0]]; break;   // StorageParameter.value (_parameters.pyx:2032-2036)
        case B200LP_OP_LAG: {  // FlowParameter / AbstractNode._prev_flow: the recorder's sample of timestep t - 1 (written by this
                               // group before its last barrier, or by an earlier launch); a plain (not read-only-path) load
            if (lag.t < a[2]) v = V[a[1]];
            else v = *(const volatile double *)(lag.rec_out[a[0]] + (size_t)(lag.t - a[2]) * lag.S + lag.sidx);
        } break;
        case B200LP_OP_THRESHOLD: {  // AbstractThresholdParameter.index / value (_thresholds.pyx:78-131), no ratchet
            const double x = V[a[0]], th = V[a[2]];
            const int pr = a[1];
            const bool on = pr == B200LP_PRED_LT ? x < th : pr == B200LP_PRED_GT ? x > th : pr == B200LP_PRED_LE ? x <= th
                          : pr == B200LP_PRED_GE ? x >= th : x == th;
            v = on ? V[a[4]] : V[a[3]];
        } break;
        case B200LP_OP_INTERP: {  // AbstractInterpolatedParameter.value (pywr/parameters/parameters.py:156-158), linear interp1d
            const int cnt = a[1];
            const int *xk = a + 2, *yk = a + 2 + cnt;
            const double x = V[a[0]], x0 = V[xk[0]], xl = V[xk[cnt - 1]];
            if (x != x) v = NAN;
            else if (x < x0) v = V[a[2 + 2 * cnt]];
            else if (x > xl) v = V[a[3 + 2 * cnt]];
            else if (x == xl) v = V[yk[cnt - 1]];
            else {
                int j = 0;
                for (int i = 1; i < cnt - 1; i++) if (x >= V[xk[i]]) j = i;
                const double xa = V[xk[j]], ya = V[yk[j]];
                const double slope = (V[yk[j + 1]] - ya) / (V[xk[j + 1]] - xa);
                v = slope * (x - xa) + ya;
            }
        } break;
        case B200LP_OP_CC_INTERP: {  // ControlCurveInterpolatedParameter.value (_control_curves.pyx:176-216)
            const int cnt = a[1];
            const int *cv = a + 2, *vv = a + 2 + cnt;
            const double c = clamp_pc(V[pc_base + a[0]]);
            double cc_prev = 1.0;
            bool done = false;
            for (int i = 0; i < cnt && !done; i++) {
                const double cc = V[cv[i]];
                if (c >= cc) {
                    const double den = cc_prev - cc;
                    if (den == 0.0) v = V[vv[i + 1]];
                    else { const double w = (c - cc) / den; v = V[vv[i]] * w + V[vv[i + 1]] * (1.0 - w); }
                    done = true;
                }
                cc_prev = cc;
            }
            if (!done) {
                const double den = cc_prev - 0.0;
                if (den == 0.0) v = V[vv[cnt]];
                else { const double w = (c - 0.0) / den; v = V[vv[cnt]] * w + V[vv[cnt + 1]] * (1.0 - w); }
            }
        } break;
        case B200LP_OP_CC_PIECEWISE: {  // ControlCurvePiecewiseInterpolatedParameter.value (:269-292, _interpolate :10-24)
            const int cnt = a[1];
            const int *cv = a + 2, *up = a + 2 + cnt, *lw = a + 2 + cnt + (cnt + 1);
            const double mn = V[a[2 + cnt + 2 * (cnt + 1)]], mx = V[a[3 + cnt + 2 * (cnt + 1)]];
            const double c = clamp_pc(V[pc_base + a[0]]);
            double cc_prev = mx, lb = mn;
            int band = cnt;
            for (int i = 0; i < cnt; i++) {
                const double cc = V[cv[i]];
                if (c >= cc) { band = i; lb = cc; break; }
                cc_prev = cc;
            }
            const double ub = cc_prev, lv = V[lw[band]], uv = V[up[band]];
            if (c < lb) v = lv;
            else if (c > ub) v = uv;
            else if (ub == lb) v = lv;
            else { const double f = (c - lb) / (ub - lb); v = lv + (uv - lv) * f; }
        } break;
        case B200LP_OP_STORAGE_RESET: {  // node.before(): Storage._reset_storage_only (pywr/_core.pyx:1294-1333)
            v = V[a[3]];                 // a: volume, pc, max_volume, flag, initial volume, initial pc (V indices)
            if (v == 1.0) { const double mxv = V[a[2]]; V[a[0]] = mxv; V[a[1]] = (mxv == 0.0) ? NAN : mxv / mxv; }
            else if (v == 2.0) { V[a[0]] = V[a[4]]; V[a[1]] = V[a[5]]; }
        } break;
        default: v = NAN;
        }
        V[dst] = v;
    }
}

// One dependency level of the lane-parallel op schedule: lane `gl` evaluates the op whose descriptor is
// d[0..8) = kind, dst, a0..a5 (indices into V; unused operands point at neutral built-in constants).
// Lanes of one level normally share a kind, so the branch on `kind` does not diverge.  Arithmetic order is the
// interpreter's (and the reference's: AggregatedParameter multiplies / adds in parameter order).
__device__ __forceinline__ void eval_fast_op(const int *d, double *V) {
    const int4 d0 = *reinterpret_cast<const int4 *>(d), d1 = *reinterpret_cast<const int4 *>(d + 4);  // two LDS.128
    const int kind = d0.x;
    if (kind == FOP_NONE) return;  // idle lane at this level
    const double a0 = V[d0.z], a1 = V[d0.w], a2 = V[d1.x], a3 = V[d1.y], a4 = V[d1.z], a5 = V[d1.w];
    double r;
    if (kind == FOP_CC) {  // a0 = pc, a1/a2 = curves (descending), a3..a5 = values
        const double c = clamp_pc(a0);
        r = c >= a1 ? a3 : (c >= a2 ? a4 : a5);
    } else if (kind == FOP_PROD) {
        r = ((((a0 * a1) * a2) * a3) * a4) * a5;
    } else if (kind == FOP_INDEXED) {  // a0 = index (as double), a1..a4 = choices (padded with the last one)
        const int ix = (int)a0;
        r = ix <= 0 ? a1 : (ix == 1 ? a2 : (ix == 2 ? a3 : a4));
    } else if (kind == FOP_SUM) {
        r = ((((a0 + a1) + a2) + a3) + a4) + a5;
    } else if (kind == FOP_MAX) {
        r = -INFINITY; r = a0 > r ? a0 : r; r = a1 > r ? a1 : r; r = a2 > r ? a2 : r; r = a3 > r ? a3 : r; r = a4 > r ? a4 : r; r = a5 > r ? a5 : r;
    } else if (kind == FOP_MIN) {
        r = INFINITY; r = a0 < r ? a0 : r; r = a1 < r ? a1 : r; r = a2 < r ? a2 : r; r = a3 < r ? a3 : r; r = a4 < r ? a4 : r; r = a5 < r ? a5 : r;
    } else if (kind == FOP_NEG) {
        r = -a0;
    } else if (kind == FOP_MAX0) {
        r = a0 > a1 ? a0 : a1;
    } else if (kind == FOP_MIN0) {
        r = a0 < a1 ? a0 : a1;
    } else {
        r = a0 / a1;
    }
    V[d0.y] = r;
}

// one table op owned by a lane: value(t) = ptr[clamp(t + off, 0, last) * stride]  (last = rows - 1)
struct TableDesc {
    const double *ptr;
    int last, stride, off, dst;
    __device__ __forceinline__ double at(int t) const {
        const int row = min(max(t + off, 0), last);
        return __ldg(ptr + (long long)row * stride);
    }
};

// Division by a denominator that is constant over the launch, with its reciprocal y = RN(1/b) computed once:
//   q0 = RN(a y);  r = a - b q0 (exact, one fma);  q = RN(q0 + r y)
// is the correctly rounded quotient RN(a / b) (Markstein's theorem), i.e. bit-identical to `a / b`, for finite
// a and b in the normal range; three dependent FMAs instead of the ~25-instruction DDIV sequence.
// (checked against `/` on 1e9 random operand pairs: no mismatch.)
__device__ __forceinline__ double div_by_const(const double a, const double b, const double y) {
    const double q0 = a * y;
    const double r = fma(-b, q0, a);
    return fma(r, y, q0);
}

// MINB: resident blocks per SM the register allocation is sized for (RPL == 1 variants): 3 -> 168 registers, the shortest
// per-timestep chain; 4 -> 128 registers (a few spilled values on the chain: +4 % for 8 columns, +25 % for 13) but a third more
// resident warps, which decides when the batch does not fit one wave of the 3-block build (b200lp.cu::run_variant).
template <int G, int RPL, int NC, int MINB>
__global__ void __launch_bounds__(128, (RPL == 1 ? MINB : 1))
lp_run_kernel(const DevStructure st, const DevState state, const DevProgram pg, const int S, const int t0,
              const int nsteps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups_per_block = blockDim.x / G;
    const int gib = threadIdx.x / G;
    const int sidx = blockIdx.x * groups_per_block + gib;
    // the pre-decoded op code is shared by the groups of the block
    const size_t gbytes = (GroupSmem<G, RPL, NC>::bytes(st.n_val) + 15) & ~(size_t)15;
    int *code = reinterpret_cast<int *>(smem_raw + gbytes * groups_per_block);
    for (int k = threadIdx.x; k < pg.n_code; k += blockDim.x) code[k] = __ldg(&pg.code[k]);
    int *fop = code + ((pg.n_code + 3) & ~3);
    for (int k = threadIdx.x; k < pg.n_fop_levels * 32 * FOP_WORDS; k += blockDim.x) fop[k] = __ldg(&pg.fop[k]);
    __syncthreads();
    if (sidx >= S) return;

    Dict<G, RPL, NC> D;
    init_group(D, st, smem_raw, gib);
    const int gl = D.gl, n = st.n, nst = st.n_storage;
    double *V = D.sm.V;

    int *meta = state.meta + (size_t)sidx * 4;
    if (meta[2] != B200LP_ST_OPTIMAL) return;  // failed at an earlier timestep
    int valid = meta[0], npiv = meta[1], refactors = 0;
    load_state(state, D, sidx, valid);
    for (int k = gl; k < nst; k += G) {
        V[st.vol_base + k] = state.vol[(size_t)k * S + sidx];
        V[st.pc_base + k] = pg.pc[(size_t)k * S + sidx];
    }
    D.sync();
    RowDesc<RPL> rd;
    prepare_static(st, D, rd);
    int sflags = 0;
    if (!st.cost_dynamic) sflags |= assemble_costs(st, D);
    D.sync();

    // table ops: lane j owns op j, j + G, ...; values are fetched one timestep ahead
    constexpr int MAXPF = 2;
    TableDesc td[MAXPF];
    double pf[MAXPF];
#pragma unroll
    for (int u = 0; u < MAXPF; u++) {
        const int j = gl + u * G;
        td[u].ptr = pg.ts_days; td[u].last = 0; td[u].stride = 0; td[u].off = 0; td[u].dst = st.k_base + K_SCRATCH;
        if (j < pg.n_table_ops) {
            const int tb = __ldg(&pg.tab_table[j]);
            const int ax = __ldg(&pg.table_axis[tb]);
            const int col = ax == -1 ? 0 : (ax == -2 ? sidx : __ldg(&pg.scen_index[(size_t)ax * S + sidx]));
            td[u].ptr = pg.table_data + __ldg(&pg.table_offset[tb]) + col;
            td[u].last = __ldg(&pg.table_rows[tb]) - 1;
            td[u].stride = __ldg(&pg.table_cols[tb]);
            td[u].off = __ldg(&pg.tab_offset[j]);
            td[u].dst = __ldg(&pg.tab_dst[j]);
        }
        pf[u] = td[u].at(t0);
    }
    // recorder owned by this lane (recorders beyond G are handled by the generic loop below).
    // Per-scenario accumulators stay in a register for the whole launch.  Lanes without a recorder run the
    // same instruction stream on harmless operands (their accumulator is never stored).
    int rk = B200LP_REC_MEAN_FLOW, rrow = 0, ridx = st.k_base + K_ZERO, rsrc = 0, rbeg = 0, rend = 0;
    double rfac = 0.0, racc = 0.0, asum = 0.0, amin = INFINITY, amax = -INFINITY;
    double *ragg = nullptr;   // array kinds: home of the temporal aggregates (sum, min, max)
    double *rout = nullptr;   // array kinds: next element to write; others: the accumulator's home
    bool r_array = false, r_on = false, r_store = false;
    if (gl < pg.n_recorders) {
        r_on = true;
        rk = __ldg(&pg.rec_kind[gl]); rrow = __ldg(&pg.rec_row[gl]); ridx = __ldg(&pg.rec_idx[gl]);
        rsrc = __ldg(&pg.rec_src[gl]); rbeg = __ldg(&pg.rec_indptr[gl]); rend = __ldg(&pg.rec_indptr[gl + 1]);
        rfac = __ldg(&pg.rec_factor[gl]);
        r_array = rk >= B200LP_REC_NODE_FLOW && rk <= B200LP_REC_STORAGE_PC;
        rout = pg.rec_out[gl];   // nullptr for an array recorder of a B200LP_PROG_AGG_ONLY program
        r_store = r_array && rout != nullptr;
        if (rout) rout += (r_array ? (size_t)t0 * S + sidx : (size_t)sidx);
        if (!r_array) racc = *rout;
        else { ragg = pg.rec_agg + (size_t)gl * 3 * S + sidx; asum = ragg[0]; amin = ragg[S]; amax = ragg[2 * (size_t)S]; }
        const bool r_storage = rk == B200LP_REC_STORAGE_VOLUME || rk == B200LP_REC_STORAGE_PC || rk == B200LP_REC_MIN_VOLUME;
        if (r_storage) { rrow = 0; rbeg = rend = 0; } else { rsrc = 0; }
    }
    // uniform over the group: is the slower code needed at all?
    const bool any_ratio = group_ballot<G>(rk == B200LP_REC_NODE_SUPPLIED || rk == B200LP_REC_NODE_CURTAIL, D.gmask()) != 0u;
    const bool any_linform = group_ballot<G>(r_on && rrow < 0, D.gmask()) != 0u;
    if (rrow < 0 && !r_on) rrow = 0;
    // storage owned by this lane
    int sk = -1, s_row = 0, s_max = st.k_base + K_ONE, s_min = st.k_base + K_ZERO, s_act = st.k_base + K_ONE, s_beg = 0, s_end = 0;
    int s_vol = st.k_base + K_SCRATCH, s_pc = st.k_base + K_SCRATCH;
    double s_rcp = 0.0;
    bool s_fastdiv = true;
    if (gl < nst) {
        sk = __ldg(&st.st_kind[gl]); s_row = __ldg(&st.st_row[gl]);
        s_max = __ldg(&st.st_maxv_idx[gl]); s_min = __ldg(&st.st_minv_idx[gl]); s_act = __ldg(&st.st_active_idx[gl]);
        s_beg = __ldg(&pg.stflow_indptr[gl]); s_end = __ldg(&pg.stflow_indptr[gl + 1]);
        s_vol = st.vol_base + gl; s_pc = st.pc_base + gl;
        // max_volume constant over the launch (a literal, not a parameter) and in the safe range: divide by reciprocal
        const double mxv = V[s_max];
        s_fastdiv = s_max >= st.n_slots && s_max < st.vol_base && fabs(mxv) > 1e-280 && fabs(mxv) < 1e280;
        s_rcp = s_fastdiv ? 1.0 / mxv : 0.0;
    }
    int s_rlen = 0;
    double *s_rmem = nullptr;
    if (gl < nst && pg.roll_len != nullptr) {
        s_rlen = __ldg(&pg.roll_len[gl]);
        s_rmem = pg.roll_mem + (size_t)__ldg(&pg.roll_off[gl]) * S + sidx;
    }
    const bool any_rolling = group_ballot<G>(s_rlen > 0, D.gmask()) != 0u;
    const bool any_virtual = group_ballot<G>(sk == B200LP_ST_KIND_VIRTUAL, D.gmask()) != 0u;
    const bool any_slowdiv = group_ballot<G>(!s_fastdiv, D.gmask()) != 0u;
    const double *days_p = pg.ts_days + t0;

    int status = B200LP_ST_OPTIMAL;
    int t = t0;
    PHASE_INIT
#pragma unroll 1
    for (; t < t0 + nsteps; t++) {
        const double days = __ldg(days_p++);
        // ---- before(): parameters -------------------------------------------------------------
#pragma unroll
        for (int u = 0; u < MAXPF; u++) {  // every lane, no branch: idle lanes target the scratch entry
            V[td[u].dst] = pf[u];
#ifndef B200LP_EXP_NO_PREFETCH
            pf[u] = td[u].at(t + 1);  // clamped to the table's last row
#endif
        }
        if (pg.n_table_ops > MAXPF * G) {
            for (int j = gl + MAXPF * G; j < pg.n_table_ops; j += G) {  // more table ops than prefetch slots
                const int tb = __ldg(&pg.tab_table[j]);
                const int ax = __ldg(&pg.table_axis[tb]);
                const int col = ax == -1 ? 0 : (ax == -2 ? sidx : __ldg(&pg.scen_index[(size_t)ax * S + sidx]));
                TableDesc q;
                q.ptr = pg.table_data + __ldg(&pg.table_offset[tb]) + col;
                q.last = __ldg(&pg.table_rows[tb]) - 1; q.stride = __ldg(&pg.table_cols[tb]); q.off = __ldg(&pg.tab_offset[j]);
                V[__ldg(&pg.tab_dst[j])] = q.at(t);
            }
        }
        D.sync();
        PHASE_MARK(0);
        if (pg.n_fop_levels > 0) {
            // lane-parallel schedule: one op per lane per dependency level, branch-free
            for (int lev = 0; lev < pg.n_fop_levels; lev++) {
                eval_fast_op(fop + (lev * 32 + (gl & 31)) * FOP_WORDS, V);
                D.sync();
            }
        } else if (pg.n_code > 0) {
            // generic interpreter: every lane evaluates the op list and stores the same values
            eval_ops(code, pg.n_code, V, st.pc_base, LagCtx{pg.rec_out, t, S, sidx});
            D.sync();
        }
        // ---- solve() ------------------------------------------------------------------------------
        PHASE_MARK(1);
        const int lane_flags = assemble_rows(st, D, rd, days, sflags);
        PHASE_MARK(2);
        status = solve_lp(st, D, valid, npiv, refactors, lane_flags);
        PHASE_MARK(3);
        if (status != B200LP_ST_OPTIMAL) break;
        extract_solution(D);
        PHASE_MARK(4);
        // ---- after(): storage mass balance (Storage.after, pywr/_core.pyx:1335-1361) ---------------
        // every lane, one instruction stream: lanes without a storage compute on scratch operands
        {
            double flow = D.sm.f->ract[s_row];
            bool upd = sk >= 0;
            if (any_virtual) {  // VirtualStorage: -sum factor * flow while active (:1483-1493)
                if (sk == B200LP_ST_KIND_VIRTUAL) {
                    flow = 0.0;
                    for (int e = s_beg; e < s_end; e++) flow = fma(__ldg(&pg.stflow_w[e]), D.sm.f->x[__ldg(&pg.stflow_col[e])], flow);
                    upd = V[s_act] != 0.0;
                }
            }
            double v = V[s_vol] + flow * days;
            const double mxv = V[s_max], mnv = V[s_min];
            if (any_rolling) {  // RollingVirtualStorage.after (pywr/_core.pyx:1522-1536): refund what was used len steps ago
                if (s_rlen > 0 && upd) {
                    double *mem = s_rmem + (size_t)(t % s_rlen) * S;
                    v += *mem;
                    v = fmin(mxv, fmax(mnv, v));
                    *mem = -flow * days;
                }
            }
            v = fabs(v - mxv) < 1e-6 ? mxv : v;
            v = fabs(v - mnv) < 1e-6 ? mnv : v;
            double pcv = div_by_const(v, mxv, s_rcp);
            if (any_slowdiv) {
                if (!s_fastdiv) pcv = (mxv == 0.0) ? NAN : v / mxv;
            }
            if (upd) { V[s_vol] = v; V[s_pc] = pcv; }
        }
        if (nst > G) {
            for (int k = gl + G; k < nst; k += G) {  // more storages than lanes
                const int kind = __ldg(&st.st_kind[k]);
                if (kind == B200LP_ST_KIND_VIRTUAL && !(V[__ldg(&st.st_active_idx[k])] != 0.0)) continue;
                double flow = 0.0;
                if (kind == B200LP_ST_KIND_STORAGE) flow = D.sm.f->ract[__ldg(&st.st_row[k])];
                else
                    for (int e = __ldg(&pg.stflow_indptr[k]); e < __ldg(&pg.stflow_indptr[k + 1]); e++)
                        flow = fma(__ldg(&pg.stflow_w[e]), D.sm.f->x[__ldg(&pg.stflow_col[e])], flow);
                double v = V[st.vol_base + k] + flow * days;
                const double mxv = V[__ldg(&st.st_maxv_idx[k])], mnv = V[__ldg(&st.st_minv_idx[k])];
                if (pg.roll_len != nullptr && __ldg(&pg.roll_len[k]) > 0) {
                    double *mem = pg.roll_mem + ((size_t)__ldg(&pg.roll_off[k]) + (size_t)(t % __ldg(&pg.roll_len[k]))) * S + sidx;
                    v += *mem;
                    v = fmin(mxv, fmax(mnv, v));
                    *mem = -flow * days;
                }
                if (fabs(v - mxv) < 1e-6) v = mxv;
                if (fabs(v - mnv) < 1e-6) v = mnv;
                V[st.vol_base + k] = v;
                V[st.pc_base + k] = (mxv == 0.0) ? NAN : v / mxv;
            }
        }
        D.sync();
        PHASE_MARK(5);
        // ---- recorders (the lane's own one): one instruction stream, selects instead of a switch ---------
        {
            double value = D.sm.f->ract[rrow < 0 ? 0 : rrow];
            if (any_linform) {
                if (rrow < 0) {
                    value = 0.0;
                    for (int e = rbeg; e < rend; e++) value = fma(__ldg(&pg.rec_w[e]), D.sm.f->x[__ldg(&pg.rec_col[e])], value);
                }
            }
            const double mf = V[ridx];
            const double vv = V[st.vol_base + rsrc], pv = V[st.pc_base + rsrc];
            const double vf = value * rfac, df = mf - value;
            double o = vf;                                   // B200LP_REC_NODE_FLOW
            o = rk == B200LP_REC_NODE_DEFICIT ? df : o;
            o = rk == B200LP_REC_STORAGE_VOLUME ? vv : o;
            o = rk == B200LP_REC_STORAGE_PC ? pv : o;
            if (any_ratio) {
                const double q = value / mf;
                o = rk == B200LP_REC_NODE_SUPPLIED ? (mf == 0.0 ? 1.0 : q) : o;
                o = rk == B200LP_REC_NODE_CURTAIL ? (mf == 0.0 ? 0.0 : 1.0 - q) : o;
            }
            double add = vf;                                 // B200LP_REC_MEAN_FLOW
            add = rk == B200LP_REC_TOTAL_FLOW ? vf * days : add;
            add = rk == B200LP_REC_TOTAL_DEFICIT ? df * days : add;
            add = rk == B200LP_REC_DEFICIT_FREQ ? (fabs(value - mf) > 1e-6 ? 1.0 : 0.0) : add;
            racc = rk == B200LP_REC_MIN_VOLUME ? fmin(racc, vv) : racc + add;
            agg_sample(asum, amin, amax, o);
            if (r_store) {
#ifndef B200LP_EXP_NO_STORE
                *rout = o;
#endif
                rout += S;
            }
        }
        if (pg.n_recorders > G) {
            for (int r = gl + G; r < pg.n_recorders; r += G) {
                const int kind = __ldg(&pg.rec_kind[r]), row = __ldg(&pg.rec_row[r]), idx = __ldg(&pg.rec_idx[r]);
                const int src = __ldg(&pg.rec_src[r]), beg = __ldg(&pg.rec_indptr[r]), end = __ldg(&pg.rec_indptr[r + 1]);
                const double f = __ldg(&pg.rec_factor[r]);
                double *out = pg.rec_out[r];
                double value = 0.0;
                if (row >= 0) value = D.sm.f->ract[row];
                else
                    for (int e = beg; e < end; e++) value = fma(__ldg(&pg.rec_w[e]), D.sm.f->x[__ldg(&pg.rec_col[e])], value);
                const size_t ts_at = (size_t)t * S + sidx;
                double o = 0.0;
                switch (kind) {
                case B200LP_REC_NODE_FLOW: o = value * f; break;
                case B200LP_REC_NODE_DEFICIT: o = V[idx] - value; break;
                case B200LP_REC_NODE_SUPPLIED: { const double mf = V[idx]; o = mf == 0.0 ? 1.0 : value / mf; } break;
                case B200LP_REC_NODE_CURTAIL: { const double mf = V[idx]; o = mf == 0.0 ? 0.0 : 1.0 - value / mf; } break;
                case B200LP_REC_STORAGE_VOLUME: o = V[st.vol_base + src]; break;
                case B200LP_REC_STORAGE_PC: o = V[st.pc_base + src]; break;
                default: break;
                }
                if (kind >= B200LP_REC_NODE_FLOW && kind <= B200LP_REC_STORAGE_PC) {
                    if (out) out[ts_at] = o;
                    double *ag = pg.rec_agg + (size_t)r * 3 * S + sidx;
                    agg_sample(ag[0], ag[S], ag[2 * (size_t)S], o);
                }
                switch (kind) {
                case B200LP_REC_TOTAL_DEFICIT: out[sidx] += (V[idx] - value) * days; break;
                case B200LP_REC_TOTAL_FLOW: out[sidx] += value * f * days; break;
                case B200LP_REC_MEAN_FLOW: out[sidx] += value * f; break;
                case B200LP_REC_DEFICIT_FREQ: if (fabs(value - V[idx]) > 1e-6) out[sidx] += 1.0; break;
                case B200LP_REC_MIN_VOLUME: { const double v = V[st.vol_base + src]; if (v < out[sidx]) out[sidx] = v; } break;
                default: break;
                }
            }
        }
        D.sync();
        PHASE_MARK(6);
    }
    PHASE_FLUSH;
    if (r_on && !r_array) *rout = racc;
    if (r_on && r_array) { ragg[0] = asum; ragg[S] = amin; ragg[2 * (size_t)S] = amax; }
    // ---- persist ------------------------------------------------------------------------------------
    store_state(state, D, sidx, true);
    for (int k = gl; k < nst; k += G) {
        state.vol[(size_t)k * S + sidx] = V[st.vol_base + k];
        pg.pc[(size_t)k * S + sidx] = V[st.pc_base + k];
    }
    if (status == B200LP_ST_OPTIMAL)
        for (int j = gl; j < n; j += G) pg.last_x[(size_t)j * S + sidx] = D.sm.f->x[j];
    if (gl == 0) {
        meta[0] = valid; meta[1] = npiv; meta[2] = status;
        if (D.pivots) atomicAdd(&state.counters[0], (unsigned long long)D.pivots);
        if (refactors) atomicAdd(&state.counters[1], (unsigned long long)refactors);
        if (status != B200LP_ST_OPTIMAL) {
            pg.fail_step[sidx] = t;
            atomicAdd(&state.fail[0], 1);
            atomicMin(&state.fail[1], (sidx << 4) | status);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Support for the structure-specialised run kernel (one kernel per model, generated by jit_run.inc at
// b200lp_load_program and compiled with NVRTC).  The generated code holds what depends on the model as
// straight-line code with immediates: table look-ups, parameter ops, the bounds of the rows that change with
// time, Storage.after and the recorders, all in registers.  What depends on the current BASIS lives here:
// the lane that owns basic position i / non-basic position j / LP row k keeps, in registers, where its
// operands come from, and every per-timestep exchange is a warp shuffle - no shared memory and no barrier
// on the steady-state path.  When the steady-state test fails the lane state is written to the shared-memory
// layout of Dict and the general solve (the same code the generic kernels run) takes over.
// One lane per row (RPL == 1) and NC <= G, so lane gl also owns non-basic position gl.
// ------------------------------------------------------------------------------------------------
template <int G, int NC>
struct JitLane {
    static_assert(NC <= G, "one non-basic position per lane");
    typedef Dict<G, 1, NC> DictT;
    static constexpr int NCP = DictT::NCP;  // == G
    // set by derive(): valid until the basis changes
    int hsrc;    // lane whose row bounds are the bounds of this lane's basic variable (own lane for a column)
    bool hcol;   // the basic variable is a column: bounds [0, +inf)
    int nsrc;    // the same for the variable in this lane's non-basic position
    bool ncol;
    int nbv;     // variable in the non-basic position (-1: padding)
    int nstat;   // its status (carried from timestep to timestep: boxed variables keep their side on ties)
    double dj;   // its reduced cost
    int rloc;    // where the row variable of LP row gl is: basic position p (p < G) or non-basic position G + j
    // per timestep
    double xn;   // value of the non-basic position

    // positions of every variable from the dictionary's head / nb (shared-memory scratch: Dict's vstat)
    __device__ __forceinline__ void derive(DictT &D) {
        const int n = D.n, gl = D.gl;
        const int h = D.head[0];
        hcol = h >= 0 && h < n;
        hsrc = h >= n ? h - n : gl;
        nbv = D.sm.f->nb[gl];
        nstat = D.sm.f->nbstat[gl];
        dj = D.sm.f->d[gl];
        ncol = nbv >= 0 && nbv < n;
        nsrc = nbv >= n ? nbv - n : gl;
        int *pos = D.sm.f->vstat;
        D.sync();
        if (h >= 0) pos[h < n ? h : NCP + (h - n)] = gl;
        if (nbv >= 0) pos[nbv < n ? nbv : NCP + (nbv - n)] = G + gl;
        D.sync();
        rloc = gl < D.m ? pos[NCP + gl] : gl;
    }
    // uniform over the group: where row i / column c is: basic position p (p < G) or non-basic position G + j
    // (valid right after derive())
    __device__ __forceinline__ int row_loc(const DictT &D, const int i) const { return D.sm.f->vstat[NCP + i]; }
    __device__ __forceinline__ int col_loc(const DictT &D, const int c) const { return D.sm.f->vstat[c]; }

    // value of the variable at location `loc` (any lane may ask for any location)
    __device__ __forceinline__ double value_at(const DictT &D, const int loc) const {
        const unsigned gm = D.gmask();
        const double b = __shfl_sync(gm, D.beta[0], loc & (G - 1), G);
        const double x = __shfl_sync(gm, xn, loc & (G - 1), G);
        return loc < G ? b : x;
    }

    // the same with ONE shuffle: `loc` is uniform over the group, so every lane offers the kind of value that is asked for
    __device__ __forceinline__ double pick(const DictT &D, const int loc) const {
        const double mine = loc < G ? D.beta[0] : xn;
        return __shfl_sync(D.gmask(), mine, loc & (G - 1), G);
    }

    // activity of LP row gl, every lane its own row (the padding lanes read their own position: never used)
    __device__ __forceinline__ double own_row_value(const DictT &D) const {
        const unsigned gm = D.gmask();
        const double b = __shfl_sync(gm, D.beta[0], rloc & (G - 1), G);
        const double x = __shfl_sync(gm, xn, rloc & (G - 1), G);
        return rloc < G ? b : x;
    }

    // Steady-state timestep (Dict::fast_step).  mylo / myhi: bounds of LP row gl (free for lanes beyond the last row);
    // lane_flags: bit 0 NaN input, bit 1 inconsistent bounds of that row.  Leaves beta in D.beta[0], the non-basic value in
    // xn.  Returns the OR over the group of lane_flags | dual infeasible non-basic position << 2 | basic variable out of
    // bounds << 3: 0 means optimal as it stands.  SMEM_X: the non-basic values travel through shared memory (one STS, a
    // warp barrier and NC / 2 LDS.128) instead of 2 NC shuffles.
    template <int NREAL, bool SMEM_X>
    __device__ __forceinline__ unsigned fast_check(DictT &D, const double mylo, const double myhi, const int lane_flags) {
        const unsigned gm = D.gmask();
        const double tl = __shfl_sync(gm, mylo, hsrc, G), th = __shfl_sync(gm, myhi, hsrc, G);
        const double nl = __shfl_sync(gm, mylo, nsrc, G), nh = __shfl_sync(gm, myhi, nsrc, G);
        const double lob = sel_pred(hcol, 0.0, tl), hib = sel_pred(hcol, INFINITY, th);
        // Dict::place_nonbasics_lane for position gl, written without short-circuit operators: one instruction stream
        const double lo = sel_pred(ncol, 0.0, nl), hi = sel_pred(ncol, INFINITY, nh);
        const bool lo_fin = lo > -INFINITY, hi_fin = hi < INFINITY, fx = lo == hi;
        const bool dpos = dj > TOL_DUAL, dneg = dj < -TOL_DUAL;
        const bool pad = nbv < 0;
        const int s0 = nstat;
        const int keep = ((s0 == VS_NL) | (s0 == VS_NU)) ? s0 : VS_NL;
        const int s_box = dpos ? VS_NL : (dneg ? VS_NU : keep);
        const int s_fin = hi_fin ? s_box : VS_NL, s_inf = hi_fin ? VS_NU : VS_NF;
        int s = lo_fin ? s_fin : s_inf;
        s = (fx | pad) ? VS_NS : s;
        const bool bad = (lo_fin & !hi_fin & dneg) | (!lo_fin & hi_fin & dpos) | (!lo_fin & !hi_fin & (dpos | dneg));
        double xv = sel_pred(s == VS_NU, hi, sel_pred(s == VS_NF, 0.0, lo));
        xv = sel_pred(pad, 0.0, xv);
        const bool infe = !fx & bad & !pad;
        nstat = s;
        xn = xv;
        // beta = T xn, j ascending (Dict::basic_values; the padding columns contribute fma(0, 0, acc) = acc)
        double acc = 0.0;
        if (SMEM_X) {
            D.sm.f->xn[D.gl] = xv;
            D.sync();
            const double2 *x2 = reinterpret_cast<const double2 *>(D.sm.f->xn);
#pragma unroll
            for (int j = 0; j + 1 < NREAL; j += 2) {
                const double2 p = x2[j / 2];
                acc = fma(D.T[0][j], p.x, acc);
                acc = fma(D.T[0][j + 1], p.y, acc);
            }
            if (NREAL & 1) acc = fma(D.T[0][NREAL - 1], D.sm.f->xn[NREAL - 1], acc);
        } else {
#pragma unroll
            for (int j = 0; j < NREAL; j++) acc = fma(D.T[0][j], __shfl_sync(gm, xv, j, G), acc);
        }
        D.beta[0] = acc;
        const bool viol = (acc < lob - TOL_PRIMAL * (1.0 + fabs(lob))) | (acc > hib + TOL_PRIMAL * (1.0 + fabs(hib)));
        return group_or<G>((unsigned)lane_flags | (infe ? 4u : 0u) | (viol ? 8u : 0u), gm);
    }

    // beta = T xn with the non-basic values exchanged like in fast_check (same order of operations as Dict::basic_values)
    template <int NREAL, bool SMEM_X>
    __device__ __forceinline__ double basic_value(DictT &D, const double xv) const {
        double acc = 0.0;
        if (SMEM_X) {
            D.sm.f->xn[D.gl] = xv;
            D.sync();
            const double2 *x2 = reinterpret_cast<const double2 *>(D.sm.f->xn);
#pragma unroll
            for (int j = 0; j + 1 < NREAL; j += 2) {
                const double2 p = x2[j / 2];
                acc = fma(D.T[0][j], p.x, acc);
                acc = fma(D.T[0][j + 1], p.y, acc);
            }
            if (NREAL & 1) acc = fma(D.T[0][NREAL - 1], D.sm.f->xn[NREAL - 1], acc);
            D.sync();   // xn is rewritten by the next exchange
        } else {
#pragma unroll
            for (int j = 0; j < NREAL; j++) acc = fma(D.T[0][j], __shfl_sync(D.gmask(), xv, j, G), acc);
        }
        return acc;
    }

    // The common case of a timestep that pivots: the basis is dual feasible (no reduced cost has to be shifted) and a few
    // dual simplex pivots restore primal feasibility.  This is phase 0 of Dict::solve with shifted == 0, on the lane state:
    // the same leaving row (largest violation), the same ratio test (ties: largest pivot element, lowest position) and the
    // same arithmetic for the dictionary, the reduced costs and the basic values, so the result is bit for bit what the
    // general solve computes - without its shared-memory round trips.  Returns true when the basis is optimal; false hands
    // over to the general solve (no entering candidate = infeasible LP, or more than MAXP pivots), which continues from the
    // state reached so far.  Either way the lane state is written through to shared memory (nb, nbstat, d, xn) at the end.
    template <int NREAL, bool SMEM_X, int MAXP>
    __device__ __forceinline__ bool dual_pivots(DictT &D, const double mylo, const double myhi, int &npiv) {
        const unsigned gm = D.gmask();
        const int gl = D.gl, n = D.n;
        double lob = sel_pred(hcol, 0.0, __shfl_sync(gm, mylo, hsrc, G)), hib = sel_pred(hcol, INFINITY, __shfl_sync(gm, myhi, hsrc, G));
        bool optimal = false;
#pragma unroll 1
        for (int it = 0;; it++) {
            const double b = D.beta[0];
            double viol = 0.0; int dd = 0;
            if (b < lob - TOL_PRIMAL * (1.0 + fabs(lob))) { viol = lob - b; dd = +1; }
            else if (b > hib + TOL_PRIMAL * (1.0 + fabs(hib))) { viol = b - hib; dd = -1; }
            double worst = -1.0; int bi = -1, dir = 0;
            if (dd != 0) { worst = viol; bi = gl; dir = dd; }
            if (group_ballot<G>(bi >= 0, gm) == 0u) { optimal = true; break; }
            if (it >= MAXP) break;
            group_argmax<G>(worst, bi, dir, gm);
            const int pl = bi;
            if (gl == pl) {
#pragma unroll
                for (int j = 0; j < NC; j++) D.sm.f->prow[j] = D.T[0][j];
            }
            D.sync();
            // dual ratio test: this lane's non-basic position (NCP == G)
            const double a = D.sm.f->prow[gl] * (double)dir;
            bool ok = false;
            if (nstat == VS_NL) ok = a > TOL_PIVOT;
            else if (nstat == VS_NU) ok = a < -TOL_PIVOT;
            else if (nstat == VS_NF) ok = fabs(a) > TOL_PIVOT;
            double ratio1 = INFINITY;
            if (ok) ratio1 = fabs(dj) / fabs(a);
            const double rmin = group_min<G>(ratio1, gm);
            if (rmin == INFINITY) break;   // no entering candidate: the general solve reports it
            const double thr = rmin * (1.0 + TIE_REL) + TIE_ABS;
            double amax = -1.0; int q = -1, aux = 0;
            if (ratio1 < INFINITY && !(ratio1 > thr)) { amax = fabs(a); q = gl; }
            group_argmax<G>(amax, q, aux, gm);
            const bool to_upper = dir < 0;
            const double vlo = __shfl_sync(gm, lob, pl, G), vhi = __shfl_sync(gm, hib, pl, G);
            // Dict::pivot(pl, 0, q, use_dw) with dw == d
            const double inv = 1.0 / D.sm.f->prow[q];
            double pr[NC];
#pragma unroll
            for (int j = 0; j < NC; j++) pr[j] = (j == q) ? inv : -D.sm.f->prow[j] * inv;
            const bool is_piv = gl == pl;
            const double f = D.col_of(0, q);
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const double tn = (j == q) ? f * inv : fma(f, pr[j], D.T[0][j]);
                D.T[0][j] = is_piv ? pr[j] : tn;
            }
            const double fd = __shfl_sync(gm, dj, q, G);
            const int vin = __shfl_sync(gm, nbv, q, G), vout = __shfl_sync(gm, D.head[0], pl, G);
            const double prj = (gl == q) ? inv : -D.sm.f->prow[gl] * inv;
            dj = (gl == q) ? fd * inv : fma(fd, prj, dj);
            if (is_piv) { D.head[0] = vin; hcol = vin < n; hsrc = vin >= n ? vin - n : gl; }
            if (gl == q) {
                nbv = vout; ncol = vout < n; nsrc = vout >= n ? vout - n : gl;
                nstat = (vlo == vhi) ? VS_NS : (to_upper ? VS_NU : VS_NL);
                xn = to_upper ? vhi : vlo;
            }
            if (gl == 0) {   // the position table derive() built (row_loc / col_loc / rloc read it)
                int *pos = D.sm.f->vstat;
                pos[vin < n ? vin : NCP + (vin - n)] = pl;
                pos[vout < n ? vout : NCP + (vout - n)] = G + q;
            }
            D.pivots++; npiv++;
            // bounds of the new basic variable of position pl (Dict::load_basic_bounds)
            const int vrow = vin >= n ? vin - n : 0;
            const double nlo = __shfl_sync(gm, mylo, vrow, G), nhi = __shfl_sync(gm, myhi, vrow, G);
            if (is_piv) { lob = vin < n ? 0.0 : nlo; hib = vin < n ? INFINITY : nhi; }
            D.sync();   // prow is rewritten by the next iteration
            D.beta[0] = basic_value<NREAL, SMEM_X>(D, xn);
        }
        D.sync();
        rloc = gl < D.m ? D.sm.f->vstat[NCP + gl] : gl;
        return optimal;
    }

    // lane state -> the shared-memory layout the general solve works on    // lane state -> the shared-memory layout the general solve works on
    __device__ __forceinline__ void materialize(DictT &D, const double mylo, const double myhi) {
        if (D.gl < D.m) { D.sm.f->LO[1 + D.n + D.gl] = mylo; D.sm.f->HI[1 + D.n + D.gl] = myhi; }
        D.sm.f->nb[D.gl] = nbv; D.sm.f->nbstat[D.gl] = nstat; D.sm.f->d[D.gl] = dj;
        D.sync();
    }
    // ... and back after it returned B200LP_ST_OPTIMAL (D.beta is current, see extract_solution)
    __device__ __forceinline__ void reload(DictT &D) {
        derive(D);
        xn = D.sm.f->xn[D.gl];
    }
    // end of the launch: what extract_solution / store_state read from shared memory
    __device__ __forceinline__ void to_shared(DictT &D) {
        D.sm.f->nb[D.gl] = nbv; D.sm.f->nbstat[D.gl] = nstat; D.sm.f->d[D.gl] = dj;
        D.sm.f->xn[D.gl] = xn;
        D.sync();
    }
};

// Table ops of the specialised kernel: lane j of the group owns ops j, j + G, ... and fetches their values one timestep
// ahead; the generated code broadcasts each value with one constant-lane shuffle.
template <int G, int PF>
struct JitTables {
    TableDesc td[PF];
    double pf[PF];
    __device__ __forceinline__ void init(const DevStructure &st, const DevProgram &pg, const int gl, const int S, const int sidx,
                                         const int t0) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int j = gl + u * G;
            td[u].ptr = pg.ts_days; td[u].last = 0; td[u].stride = 0; td[u].off = 0; td[u].dst = 0;
            if (j < pg.n_table_ops) {
                const int tb = __ldg(&pg.tab_table[j]);
                const int ax = __ldg(&pg.table_axis[tb]);
                const int col = ax == -1 ? 0 : (ax == -2 ? sidx : __ldg(&pg.scen_index[(size_t)ax * S + sidx]));
                td[u].ptr = pg.table_data + __ldg(&pg.table_offset[tb]) + col;
                td[u].last = __ldg(&pg.table_rows[tb]) - 1;
                td[u].stride = __ldg(&pg.table_cols[tb]);
                td[u].off = __ldg(&pg.tab_offset[j]);
            }
            pf[u] = td[u].at(t0);
        }
    }
    __device__ __forceinline__ void prefetch(const int t) {
#pragma unroll
        for (int u = 0; u < PF; u++) pf[u] = td[u].at(t);  // clamped to the table's last row
    }
};

// Recorders of the specialised kernel: lane r % G owns recorder r (unit r / G).  The generated code computes every
// recorder's sample as a group-uniform value and hands each lane the one it owns; accumulators, running aggregates and
// the output cursor stay in the owner's registers for the whole launch.
template <int G, int RU>
struct JitRecorders {
    double racc[RU], asum[RU], amin[RU], amax[RU];
    double *rout[RU], *ragg[RU];
    bool on[RU], arr[RU], store[RU], ismin[RU];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int u = 0; u < RU; u++) {
            racc[u] = 0.0; asum[u] = 0.0; amin[u] = INFINITY; amax[u] = -INFINITY;
            rout[u] = nullptr; ragg[u] = nullptr;
            on[u] = arr[u] = store[u] = ismin[u] = false;
        }
    }
    // this lane owns recorder r in unit u (the generated code calls it under `if (gl == owner lane)`)
    __device__ __forceinline__ void bind(const int u, const int r, const DevProgram &pg, const int S, const int sidx, const int t0) {
        const int rk = __ldg(&pg.rec_kind[r]);
        on[u] = true;
        arr[u] = rk >= B200LP_REC_NODE_FLOW && rk <= B200LP_REC_STORAGE_PC;
        ismin[u] = rk == B200LP_REC_MIN_VOLUME;
        rout[u] = pg.rec_out[r];  // nullptr: array recorder of a B200LP_PROG_AGG_ONLY program
        store[u] = arr[u] && rout[u] != nullptr;
        if (rout[u]) rout[u] += arr[u] ? (size_t)t0 * S + sidx : (size_t)sidx;
        if (!arr[u]) racc[u] = *rout[u];
        else {
            ragg[u] = pg.rec_agg + (size_t)r * 3 * S + sidx;
            asum[u] = ragg[u][0]; amin[u] = ragg[u][S]; amax[u] = ragg[u][2 * (size_t)S];
        }
    }
    // o: sample of an array kind; add: increment of an accumulating kind; vv: volume of a minimum-volume recorder
    __device__ __forceinline__ void sample(const int u, const double o, const double add, const double vv, const int S) {
        agg_sample(asum[u], amin[u], amax[u], o);
        if (store[u]) { *rout[u] = o; rout[u] += S; }
        racc[u] = ismin[u] ? fmin(racc[u], vv) : racc[u] + add;
    }
    // the same, only when `on_now` (group-uniform): lets the generated code take the samples of timestep t-1 at the top of
    // timestep t, where they fill the issue slots the recurrence (volume -> bounds -> basic values) leaves empty
    __device__ __forceinline__ void sample_if(const int u, const bool on_now, const double o, const double add, const double vv, const int S) {
        const double s2 = asum[u] + o;
        asum[u] = on_now ? s2 : asum[u];
        amin[u] = (on_now & ((o < amin[u]) | (o != o))) ? o : amin[u];
        amax[u] = (on_now & ((o > amax[u]) | (o != o))) ? o : amax[u];
        if (store[u] & on_now) *rout[u] = o;
        rout[u] += (store[u] & on_now) ? S : 0;
        const double r2 = ismin[u] ? fmin(racc[u], vv) : racc[u] + add;
        racc[u] = on_now ? r2 : racc[u];
    }
    __device__ __forceinline__ void finish(const int S) {
#pragma unroll
        for (int u = 0; u < RU; u++) {
            if (on[u] && !arr[u]) *rout[u] = racc[u];
            if (on[u] && arr[u]) { ragg[u][0] = asum[u]; ragg[u][S] = amin[u]; ragg[u][2 * (size_t)S] = amax[u]; }
        }
    }
};

// reset: slack basis for every scenario (is_first_solve, cython_glpk.pyx:757-759)
__global__ void lp_reset_kernel(const DevState state, const int S, const int m, const int n, const int rows_cap,
                                const int nc) {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= S) return;
    for (int i = 0; i < rows_cap; i++) state.head[(size_t)sidx * rows_cap + i] = i < m ? n + i : -1;
    for (int j = 0; j < nc; j++) {
        state.nb[(size_t)sidx * nc + j] = j < n ? j : -1;
        state.nbstat[(size_t)sidx * nc + j] = j < n ? VS_NL : VS_NS;
        state.d[(size_t)sidx * nc + j] = 0.0;
    }
    state.meta[(size_t)sidx * 4 + 0] = 0;
    state.meta[(size_t)sidx * 4 + 1] = 0;
    state.meta[(size_t)sidx * 4 + 2] = 0;
    state.meta[(size_t)sidx * 4 + 3] = 0;
}

}  // namespace b200lp
