// rsx_kernels.cuh — device instantiation of rsx_core.cuh: one CTA of RSX_THREADS threads per scenario, working
// vectors in dynamic shared memory (<= 227 KB), the scenario's basis inverse W streamed from HBM.  CTAs are
// persistent (one per SM, shared memory allows no more) and pull scenarios from a global counter, because the
// number of pivots differs from scenario to scenario.
#pragma once
#include <cuda_runtime.h>

#include "rsx_core.cuh"

namespace b200lp {

constexpr int RSX_THREADS = 512;
// threads of a block when two blocks share an SM (MINB = 2): 512 leaves 64 registers per thread, 384 leaves 85
#ifndef RSX_THREADS2
#define RSX_THREADS2 512
#endif
constexpr int rsx_threads(int minb) { return minb == 2 ? RSX_THREADS2 : RSX_THREADS; }

struct RsxCtx {
    static constexpr int nl = 32;  // compile-time: the streaming loops unroll over it
    int tid, nt, lane, warp, nw;
    double *rv;  // [32] reduction scratch (static shared memory)
    int *ri, *ra;
    // phase timing of debug builds (-DRSX_PHASE_TIMING): cycles of thread 0 between marks, summed per phase
#ifdef RSX_PHASE_TIMING
    long long t_last;
    unsigned long long ph[16];
    __device__ __forceinline__ void mark(int k) {
        const long long c = clock64();
        if (k > 0) ph[k] += (unsigned long long)(c - t_last);
        t_last = c;
    }
#else
    __device__ __forceinline__ void mark(int) {}
#endif
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ bool leader() const { return tid == 0; }
    __device__ __forceinline__ int atomic_inc(int *p) { return atomicAdd(p, 1); }
    __device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    }
    __device__ __forceinline__ int bor(int v) { return __syncthreads_or(v); }
    // arg-max of (key, lowest idx on ties) over a warp: keys of candidates are > 0 (bound violations, |pivot elements|, 1.0 under
    // Bland's rule), so their bit patterns order like the numbers - two REDUX for the largest key, one for the lowest index
    // among the lanes that hold it, the winner's data by shuffle (lp_kernel.cuh::group_argmax)
    __device__ __forceinline__ void warp_argmax(double &key, int &idx, int &aux) {
        const bool cand = idx >= 0;
        const unsigned long long b = cand ? (unsigned long long)__double_as_longlong(key) + 1ull : 0ull;
        const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        const bool top = cand && hi == mh && lo == ml;
        const int best = (int)__reduce_min_sync(0xffffffffu, top ? (unsigned)idx : 0xffffffffu);
        if (best < 0) { idx = -1; return; }
        const int src = __ffs(__ballot_sync(0xffffffffu, top && idx == best)) - 1;
        key = __shfl_sync(0xffffffffu, key, src);
        aux = __shfl_sync(0xffffffffu, aux, src);
        idx = best;
    }
    // arg-max of (key, lowest idx on ties) over the thread-local candidates; idx < 0: no candidate
    __device__ __forceinline__ void argmax(double &key, int &idx, int &aux) {
        warp_argmax(key, idx, aux);
        if (lane == 0) { rv[warp] = key; ri[warp] = idx; ra[warp] = aux; }
        __syncthreads();
        // every warp reduces the per-warp winners again (at most 32 warps): no serial pass
        key = lane < nw ? rv[lane] : -1.0; idx = lane < nw ? ri[lane] : -1; aux = lane < nw ? ra[lane] : 0;
        warp_argmax(key, idx, aux);
        __syncthreads();
    }
    // minimum of values that are >= 0 or +inf (ratios)
    __device__ __forceinline__ double warp_min(double v) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);   // -0.0 -> +0.0
        const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
        const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
        return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
    }
    __device__ __forceinline__ double min(double v) {
        v = warp_min(v);
        if (lane == 0) rv[warp] = v;
        __syncthreads();
        v = warp_min(lane < nw ? rv[lane] : INFINITY);
        __syncthreads();
        return v;
    }
};

struct RsxDev {
    rsx::Lp lp;
    double *W;            // [S][m][ldw]
    int *head;            // [S][m]
    unsigned char *stat;  // [S][n+m]
    double *d;            // [S][n+m]
    int *info;            // [S][4]
    const double *c, *lc, *uc, *lo, *hi;  // planes of the PDHG front end, [item][ld]
    double *x;
    int ld, S;
    int *status;          // [ld] preset by the assembly (NaN / inconsistent bounds), result of the solve
    int *queue;           // [1] next scenario
    unsigned long long *counters;  // [0] pivots, [1] refactorisations
    double *bt;           // [S][2 (n+m)] bounds of every scenario, contiguous per scenario (rsx_gather_bounds)
    double *xt;           // [S][n] column values, contiguous per scenario (rsx_scatter_x moves them to the x planes)
    unsigned char *scratch;        // MODE 1: [gridDim.x][scratch_stride] reduced costs / pivot row / index lists of each block
    size_t scratch_stride;
};

// MODE 0: all working vectors in shared memory.  MODE 1: reduced costs, pivot row and index lists in an L2-resident scratch
// (rsx_core.cuh::work_bytes), which halves the shared-memory footprint.  That matters for the STREAM, not for occupancy: with
// more than ~164 KB of shared memory carved out of the SM's 256 KB, what is left as L1 cannot hold the sectors in flight and the
// same loop drops from 7.1 to 5.2 TB/s (benchmarks/tools/stream_probe.cu, profiles/r1_stream_probe.md).  MINB = blocks per SM
// (128 or 64 registers per thread).
template <int MODE, int MINB>
__global__ void __launch_bounds__(rsx_threads(MINB), MINB)
rsx_solve_kernel(const RsxDev D, const int s0, const int s1, const int cold) {
    extern __shared__ __align__(16) unsigned char rsx_smem[];
    __shared__ double rv[32];
    __shared__ int ri[32], ra[32], s_next;
    RsxCtx cx;
    cx.tid = threadIdx.x; cx.nt = blockDim.x; cx.lane = threadIdx.x & 31;
    cx.warp = threadIdx.x >> 5; cx.nw = blockDim.x >> 5;
    cx.rv = rv; cx.ri = ri; cx.ra = ra;
#ifdef RSX_PHASE_TIMING
    for (int k = 0; k < 16; k++) cx.ph[k] = 0;
    cx.t_last = clock64();
#endif
    const rsx::Lp &P = D.lp;
    rsx::Work K;
    rsx::work_carve(K, rsx_smem, MODE == 0 ? nullptr : D.scratch + (size_t)blockIdx.x * D.scratch_stride, P.m, P.n, MODE);
    const size_t wtot = (size_t)P.m * P.ldw;
    const int nv = P.n + P.m;
    int pivots = 0, refactors = 0;
    for (;;) {
        if (threadIdx.x == 0) s_next = s0 + atomicAdd(D.queue, 1);
        __syncthreads();
        const int s = s_next;
        __syncthreads();
        if (s >= s1) break;
        if (D.status[s] != 0) continue;
        rsx::Scen sc;
        sc.W = D.W + (size_t)s * wtot;
        sc.head = D.head + (size_t)s * P.m;
        sc.stat = D.stat + (size_t)s * nv;
        sc.d = D.d + (size_t)s * nv;
        sc.info = D.info + (size_t)s * 4;
        sc.c = D.c + s; sc.lc = D.lc + s; sc.uc = D.uc + s; sc.lo = D.lo + s; sc.hi = D.hi + s;
        sc.ld = (size_t)D.ld;
        sc.bt = D.bt + (size_t)s * 2 * nv;
        sc.x = D.xt + (size_t)s * P.n; sc.ldx = 1;
        const int st = rsx::run_scenario(cx, P, sc, K, cold, pivots, refactors);
        if (threadIdx.x == 0) D.status[s] = st;
    }
    if (threadIdx.x == 0) {
        if (pivots) atomicAdd(&D.counters[0], (unsigned long long)pivots);
        if (refactors) atomicAdd(&D.counters[1], (unsigned long long)refactors);
#ifdef RSX_PHASE_TIMING
        for (int k = 0; k < 16; k++) atomicAdd(&D.counters[2 + k], cx.ph[k]);
#endif
    }
}

// Crossover of the PDHG path (rsx_core.cuh::crossover): same CTA layout as rsx_solve_kernel; x0 = the scenario's column of the
// x planes (the first-order method's answer), the vertex goes to xt.  counters[20] counts the structurals the crash pivoted in.
template <int MODE, int MINB>
__global__ void __launch_bounds__(rsx_threads(MINB), MINB)
rsx_crossover_kernel(const RsxDev D, const int s0, const int s1) {
    extern __shared__ __align__(16) unsigned char rsx_smem[];
    __shared__ double rv[32];
    __shared__ int ri[32], ra[32], s_next;
    RsxCtx cx;
    cx.tid = threadIdx.x; cx.nt = blockDim.x; cx.lane = threadIdx.x & 31;
    cx.warp = threadIdx.x >> 5; cx.nw = blockDim.x >> 5;
    cx.rv = rv; cx.ri = ri; cx.ra = ra;
#ifdef RSX_PHASE_TIMING
    for (int k = 0; k < 16; k++) cx.ph[k] = 0;
    cx.t_last = clock64();
#endif
    const rsx::Lp &P = D.lp;
    rsx::Work K;
    rsx::work_carve(K, rsx_smem, MODE == 0 ? nullptr : D.scratch + (size_t)blockIdx.x * D.scratch_stride, P.m, P.n, MODE);
    const size_t wtot = (size_t)P.m * P.ldw;
    const int nv = P.n + P.m;
    int pivots = 0, refactors = 0, crashed = 0;
    for (;;) {
        if (threadIdx.x == 0) s_next = s0 + atomicAdd(D.queue, 1);
        __syncthreads();
        const int s = s_next;
        __syncthreads();
        if (s >= s1) break;
        if (D.status[s] != 0) continue;
        rsx::Scen sc;
        sc.W = D.W + (size_t)s * wtot;
        sc.head = D.head + (size_t)s * P.m;
        sc.stat = D.stat + (size_t)s * nv;
        sc.d = D.d + (size_t)s * nv;
        sc.info = D.info + (size_t)s * 4;
        sc.c = D.c + s; sc.lc = D.lc + s; sc.uc = D.uc + s; sc.lo = D.lo + s; sc.hi = D.hi + s;
        sc.ld = (size_t)D.ld;
        sc.bt = D.bt + (size_t)s * 2 * nv;
        sc.x = D.xt + (size_t)s * P.n; sc.ldx = 1;
        const int st = rsx::crossover(cx, P, sc, K, D.x + s, (size_t)D.ld, pivots, refactors, crashed);
        if (threadIdx.x == 0) D.status[s] = st;
    }
    if (threadIdx.x == 0) {
        if (pivots) atomicAdd(&D.counters[0], (unsigned long long)pivots);
        if (refactors) atomicAdd(&D.counters[1], (unsigned long long)refactors);
        if (crashed) atomicAdd(&D.counters[20], (unsigned long long)crashed);
    }
}

// Scenarios the first-order iterations left at their iteration limit go through the crossover as well: the simplex finishes from
// whatever point they reached (and reports infeasibility where that was the reason), instead of the run failing on them.
__global__ void rsx_requeue_unconverged(int *__restrict__ status, const int S) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S && status[s] == B200LP_ST_ITERLIMIT) status[s] = B200LP_ST_OPTIMAL;
}

// planes [item][ld] -> per-scenario vectors [S][2 (n+m)] (LO | HI), 32 x 32 tiles through shared memory: coalesced reads along
// the scenario axis, coalesced writes along the item axis
__global__ void rsx_gather_bounds(const RsxDev D) {
    __shared__ double tile[32][33];
    const int n = D.lp.n, nv = D.lp.n + D.lp.m, np = 2 * nv;
    const int p0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, s = s0 + tx;
        double v = 0.0;
        if (p < np && s < D.S) {
            const int q = p < nv ? p : p - nv;
            const double *plane = p < nv ? (q < n ? D.lc + (size_t)q * D.ld : D.lo + (size_t)(q - n) * D.ld)
                                         : (q < n ? D.uc + (size_t)q * D.ld : D.hi + (size_t)(q - n) * D.ld);
            v = plane[s];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r, p = p0 + tx;
        if (s < D.S && p < np) D.bt[(size_t)s * np + p] = tile[tx][r];
    }
}

// per-scenario x [S][n] -> planes [n][ld]
__global__ void rsx_scatter_x(const RsxDev D) {
    __shared__ double tile[32][33];
    const int n = D.lp.n;
    const int j0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r, j = j0 + tx;
        tile[r][tx] = (s < D.S && j < n) ? D.xt[(size_t)s * n + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, s = s0 + tx;
        if (j < n && s < D.S && D.status[s] == 0) D.x[(size_t)j * D.ld + s] = tile[tx][r];
    }
}

// Basis of every scenario in the ORIGINAL row space, GLPK-coded (BasisManager.row_stat / col_stat, cython_glpk.pyx:197-232):
// a kept row carries the status of its slack; a dropped row is basic; a column that is non-basic in the reduced LP at a bound
// that came from one of its singleton rows (pdhg_assemble_cols) is, in the original LP, basic with THAT row non-basic at the
// bound - the same vertex, the same number (m) of basic variables.  One thread per (scenario, column); the singleton rows of
// different columns are disjoint.
__global__ void rsx_basis_kernel(const RsxDev D, const PdhgDev P, const double *__restrict__ slots, const double *__restrict__ vol,
                                 const double days, int *__restrict__ row_stat, int *__restrict__ col_stat) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (s >= D.S) return;
    const int n = P.n, m = P.m, nv = n + D.lp.m;
    const unsigned char *stat = D.stat + (size_t)s * nv;
    int *rs = row_stat + (size_t)s * m, *cs = col_stat + (size_t)s * n;
    if (j == 0) {  // this thread also maps the kept rows (after every thread of the scenario wrote "basic" below: disjoint rows)
        for (int r = 0; r < D.lp.m; r++) rs[__ldg(&P.keep[r])] = stat[n + r];
    }
    if (__ldg(&P.alias_root[j]) != j) {   // folded column (doubleton elimination): basic, its row an equality on its bound
        rs[__ldg(&P.elim_row[j])] = rsx::V_NS;
        cs[j] = rsx::V_BASIC;
        return;
    }
    int st = stat[j];
    double lc = 0.0, uc = INFINITY;
    int src_lo = -1, src_hi = -1, code_lo = 0, code_hi = 0;
    for (int e = __ldg(&P.sing_ptr[j]); e < __ldg(&P.sing_ptr[j + 1]); e++) {
        const int i = __ldg(&P.sing_row[e]);
        double l, u;
        pd_row_bounds(P, slots, vol, i, s, days, l, u);
        const double a = __ldg(&P.sing_coef[e]);
        const double l2 = a > 0.0 ? l / a : u / a, u2 = a > 0.0 ? u / a : l / a;
        if (l2 > lc) { lc = l2; src_lo = i; code_lo = (l == u) ? rsx::V_NS : (a > 0.0 ? rsx::V_NL : rsx::V_NU); }
        if (u2 < uc) { uc = u2; src_hi = i; code_hi = (l == u) ? rsx::V_NS : (a > 0.0 ? rsx::V_NU : rsx::V_NL); }
        rs[i] = rsx::V_BASIC;
    }
    if (st == rsx::V_NS) st = src_lo >= 0 ? rsx::V_NL : (src_hi >= 0 ? rsx::V_NU : st);
    if (st == rsx::V_NL && src_lo >= 0) { rs[src_lo] = code_lo; st = rsx::V_BASIC; }
    else if (st == rsx::V_NU && src_hi >= 0) { rs[src_hi] = code_hi; st = rsx::V_BASIC; }
    cs[j] = st;
}

// basis, reduced costs and inverse of scenario `src` become the starting point of every other scenario
__global__ void rsx_seed_kernel(const RsxDev D, const int src, const int s_off) {
    const int s = s_off + blockIdx.y;
    if (s == src) return;
    const rsx::Lp &P = D.lp;
    const size_t wtot2 = (size_t)P.m * P.ldw / 2;  // ldw is even and the arrays are 16-byte aligned
    const double2 *from = reinterpret_cast<const double2 *>(D.W + (size_t)src * wtot2 * 2);
    double2 *to = reinterpret_cast<double2 *>(D.W + (size_t)s * wtot2 * 2);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < wtot2; e += (size_t)gridDim.x * blockDim.x) to[e] = from[e];
    if (blockIdx.x == 0) {
        const int nv = P.n + P.m;
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            D.stat[(size_t)s * nv + v] = D.stat[(size_t)src * nv + v];
            D.d[(size_t)s * nv + v] = D.d[(size_t)src * nv + v];
        }
        for (int i = threadIdx.x; i < P.m; i += blockDim.x) D.head[(size_t)s * P.m + i] = D.head[(size_t)src * P.m + i];
        if (threadIdx.x < 4) D.info[(size_t)s * 4 + threadIdx.x] = D.info[(size_t)src * 4 + threadIdx.x];
    }
}

}  // namespace b200lp
