// rsx_kernels.cuh — device instantiation of rsx_core.cuh: one CTA of RSX_THREADS threads per scenario, working
// vectors in dynamic shared memory (<= 227 KB), the scenario's basis inverse W streamed from HBM.  CTAs are
// persistent (one per SM, shared memory allows no more) and pull scenarios from a global counter, because the
// number of pivots differs from scenario to scenario.
#pragma once
#include <cuda_runtime.h>

#include "rsx_core.cuh"

namespace b200lp {

constexpr int RSX_THREADS = 512;

struct RsxCtx {
    int tid, nt, lane, nl, warp, nw;
    double *rv;  // [32] reduction scratch (static shared memory)
    int *ri, *ra;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ bool leader() const { return tid == 0; }
    __device__ __forceinline__ int atomic_inc(int *p) { return atomicAdd(p, 1); }
    __device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    }
    __device__ __forceinline__ int bor(int v) { return __syncthreads_or(v); }
    // arg-max of (key, lowest idx on ties) over the thread-local candidates; idx < 0: no candidate
    __device__ __forceinline__ void argmax(double &key, int &idx, int &aux) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double k2 = __shfl_xor_sync(0xffffffffu, key, off);
            const int i2 = __shfl_xor_sync(0xffffffffu, idx, off);
            const int a2 = __shfl_xor_sync(0xffffffffu, aux, off);
            const bool take = (i2 >= 0) && (idx < 0 || k2 > key || (k2 == key && i2 < idx));
            if (take) { key = k2; idx = i2; aux = a2; }
        }
        if (lane == 0) { rv[warp] = key; ri[warp] = idx; ra[warp] = aux; }
        __syncthreads();
        key = rv[0]; idx = ri[0]; aux = ra[0];
        for (int w = 1; w < nw; w++) {
            const double k2 = rv[w];
            const int i2 = ri[w], a2 = ra[w];
            const bool take = (i2 >= 0) && (idx < 0 || k2 > key || (k2 == key && i2 < idx));
            if (take) { key = k2; idx = i2; aux = a2; }
        }
        __syncthreads();
    }
    __device__ __forceinline__ double min(double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) rv[warp] = v;
        __syncthreads();
        v = rv[0];
        for (int w = 1; w < nw; w++) v = fmin(v, rv[w]);
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
};

__global__ void __launch_bounds__(RSX_THREADS, 1)
rsx_solve_kernel(const RsxDev D, const int s0, const int s1, const int cold) {
    extern __shared__ __align__(16) unsigned char rsx_smem[];
    __shared__ double rv[32];
    __shared__ int ri[32], ra[32], s_next;
    RsxCtx cx;
    cx.tid = threadIdx.x; cx.nt = blockDim.x; cx.lane = threadIdx.x & 31; cx.nl = 32;
    cx.warp = threadIdx.x >> 5; cx.nw = blockDim.x >> 5;
    cx.rv = rv; cx.ri = ri; cx.ra = ra;
    const rsx::Lp &P = D.lp;
    rsx::Work K;
    rsx::work_carve(K, rsx_smem, P.m, P.n);
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
        sc.c = D.c + s; sc.lc = D.lc + s; sc.uc = D.uc + s; sc.lo = D.lo + s; sc.hi = D.hi + s; sc.x = D.x + s;
        sc.ld = (size_t)D.ld;
        const int st = rsx::run_scenario(cx, P, sc, K, cold, pivots, refactors);
        if (threadIdx.x == 0) D.status[s] = st;
    }
    if (threadIdx.x == 0) {
        if (pivots) atomicAdd(&D.counters[0], (unsigned long long)pivots);
        if (refactors) atomicAdd(&D.counters[1], (unsigned long long)refactors);
    }
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
