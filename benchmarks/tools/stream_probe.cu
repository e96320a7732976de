// stream_probe.cu — how fast can one CTA per SM stream "its" 12 MB basis inverse (row-major m x ldw doubles) out of HBM
// while forming beta = W g?  Variants of the inner loop of rsx_core.cuh::compute_beta, measured on S scenarios.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_probe stream_probe.cu && ./stream_probe [S] [m]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int POL> __device__ __forceinline__ double ldp(const double *p) {
    if (POL == 1) return __ldcg(p);
    if (POL == 2) return __ldcs(p);
    if (POL == 3) return __ldlu(p);
    if (POL == 4) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
    return *p;
}
// V0P: v0 with a cache policy on the stream
template <int POL>
__global__ void __launch_bounds__(512, 1) v0p(const double *W, const double *g_in, double *out, int S, int m, int ldw, int *queue) {
    extern __shared__ double g[];
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < m; k += blockDim.x) g[k] = g_in[k];
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int s = s_next;
        __syncthreads();
        if (s >= S) break;
        const double *Ws = W + (size_t)s * m * ldw;
        for (int i = 2 * warp; i < m; i += 2 * nw) {
            const double *r0 = Ws + (size_t)i * ldw, *r1 = i + 1 < m ? r0 + ldw : r0;
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
            int k = lane;
#pragma unroll 4
            for (; k + 96 < m; k += 128) {
                const double x0 = ldp<POL>(r0 + k), x1 = ldp<POL>(r0 + k + 32), x2 = ldp<POL>(r0 + k + 64), x3 = ldp<POL>(r0 + k + 96);
                const double y0 = ldp<POL>(r1 + k), y1 = ldp<POL>(r1 + k + 32), y2 = ldp<POL>(r1 + k + 64), y3 = ldp<POL>(r1 + k + 96);
                a0 = fma(x0, g[k], a0); a1 = fma(x1, g[k + 32], a1); a2 = fma(x2, g[k + 64], a2); a3 = fma(x3, g[k + 96], a3);
                b0 = fma(y0, g[k], b0); b1 = fma(y1, g[k + 32], b1); b2 = fma(y2, g[k + 64], b2); b3 = fma(y3, g[k + 96], b3);
            }
            for (; k < m; k += 32) { a0 = fma(r0[k], g[k], a0); b0 = fma(r1[k], g[k], b0); }
            const double sa = warp_sum((a0 + a1) + (a2 + a3)), sb = warp_sum((b0 + b1) + (b2 + b3));
            if (lane == 0) { out[(size_t)s * m + i] = sa; if (i + 1 < m) out[(size_t)s * m + i + 1] = sb; }
        }
    }
}

// V0: the loop of the product kernel: two rows per warp, four 256-byte segments per row
__global__ void __launch_bounds__(512, 1) v0(const double *W, const double *g_in, double *out, int S, int m, int ldw, int *queue) {
    extern __shared__ double g[];
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < m; k += blockDim.x) g[k] = g_in[k];
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int s = s_next;
        __syncthreads();
        if (s >= S) break;
        const double *Ws = W + (size_t)s * m * ldw;
        for (int i = 2 * warp; i < m; i += 2 * nw) {
            const double *r0 = Ws + (size_t)i * ldw, *r1 = i + 1 < m ? r0 + ldw : r0;
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
            int k = lane;
            for (; k + 96 < m; k += 128) {
                const double x0 = r0[k], x1 = r0[k + 32], x2 = r0[k + 64], x3 = r0[k + 96];
                const double y0 = r1[k], y1 = r1[k + 32], y2 = r1[k + 64], y3 = r1[k + 96];
                a0 = fma(x0, g[k], a0); a1 = fma(x1, g[k + 32], a1); a2 = fma(x2, g[k + 64], a2); a3 = fma(x3, g[k + 96], a3);
                b0 = fma(y0, g[k], b0); b1 = fma(y1, g[k + 32], b1); b2 = fma(y2, g[k + 64], b2); b3 = fma(y3, g[k + 96], b3);
            }
            for (; k < m; k += 32) { a0 = fma(r0[k], g[k], a0); b0 = fma(r1[k], g[k], b0); }
            const double sa = warp_sum((a0 + a1) + (a2 + a3)), sb = warp_sum((b0 + b1) + (b2 + b3));
            if (lane == 0) { out[(size_t)s * m + i] = sa; if (i + 1 < m) out[(size_t)s * m + i + 1] = sb; }
        }
    }
}

// V1: 128-bit loads, the whole matrix of the scenario as one flat stream: thread t handles double2 elements t, t+nt, ...
// of NROWS consecutive rows at a time (row-major, ldw even), UNR independent loads in flight per thread
template <int UNR, int MINB>
__global__ void __launch_bounds__(512, MINB) v1(const double *W, const double *g_in, double *out, int S, int m, int ldw, int *queue) {
    extern __shared__ double g[];
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < ldw; k += blockDim.x) g[k] = k < m ? g_in[k] : 0.0;
    const int l2 = ldw / 2;
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int s = s_next;
        __syncthreads();
        if (s >= S) break;
        const double2 *Ws = reinterpret_cast<const double2 *>(W + (size_t)s * m * ldw);
        const double2 *g2 = reinterpret_cast<const double2 *>(g);
        for (int i = warp; i < m; i += nw) {
            const double2 *row = Ws + (size_t)i * l2;
            double acc[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) acc[u] = 0.0;
            int k = lane;
            for (; k + (UNR - 1) * 32 < l2; k += UNR * 32) {
                double2 x[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) x[u] = __ldcs(row + k + u * 32);
#pragma unroll
                for (int u = 0; u < UNR; u++) { const double2 gg = g2[k + u * 32]; acc[u] = fma(x[u].x, gg.x, fma(x[u].y, gg.y, acc[u])); }
            }
            for (; k < l2; k += 32) { const double2 x = __ldcs(row + k); const double2 gg = g2[k]; acc[0] = fma(x.x, gg.x, fma(x.y, gg.y, acc[0])); }
            double sa = 0.0;
#pragma unroll
            for (int u = 0; u < UNR; u++) sa += acc[u];
            sa = warp_sum(sa);
            if (lane == 0) out[(size_t)s * m + i] = sa;
        }
    }
}

// V3: rows pulled into a shared-memory ring by 1-D bulk copies (cp.async.bulk + mbarrier), one elected thread issues,
// all warps consume: STAGES chunks of CH rows in flight per CTA
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}

template <int STAGES, int CH>
__global__ void __launch_bounds__(512, 1) v3(const double *W, const double *g_in, double *out, int S, int m, int ldw, int *queue) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_next;
    __shared__ __align__(8) uint64_t full[STAGES];
    double *g = reinterpret_cast<double *>(smem);
    double *ring = g + ((ldw + 15) & ~15);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < ldw; k += blockDim.x) g[k] = k < m ? g_in[k] : 0.0;
    if (threadIdx.x == 0) { for (int q = 0; q < STAGES; q++) mbar_init(&full[q], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t chunk_elems = (size_t)CH * ldw;
    uint32_t phase_bits = 0;  // per-stage parity
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        const int s = s_next;
        __syncthreads();
        if (s >= S) break;
        const double *Ws = W + (size_t)s * m * ldw;
        const int nchunks = (m + CH - 1) / CH;
        auto issue = [&](int c) {
            const int st = c % STAGES;
            const int rows = min(CH, m - c * CH);
            const uint32_t bytes = (uint32_t)(rows * ldw * sizeof(double));
            mbar_expect(&full[st], bytes);
            bulk_g2s(ring + (size_t)st * chunk_elems, Ws + (size_t)c * chunk_elems, bytes, &full[st]);
        };
        if (threadIdx.x == 0) for (int c = 0; c < STAGES - 1 && c < nchunks; c++) issue(c);
        for (int c = 0; c < nchunks; c++) {
            const int st = c % STAGES;
            // the stage that chunk c + STAGES - 1 will use was consumed in iteration c - 1 (all threads passed the barrier below)
            if (threadIdx.x == 0 && c + STAGES - 1 < nchunks) issue(c + STAGES - 1);
            mbar_wait(&full[st], (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
            const double *buf = ring + (size_t)st * chunk_elems;
            const int rows = min(CH, m - c * CH);
            for (int r = warp; r < rows; r += nw) {
                const double2 *row = reinterpret_cast<const double2 *>(buf + (size_t)r * ldw);
                const double2 *g2 = reinterpret_cast<const double2 *>(g);
                double a0 = 0.0, a1 = 0.0;
                int k = lane;
                for (; k + 32 < ldw / 2; k += 64) {
                    const double2 x0 = row[k], x1 = row[k + 32], h0 = g2[k], h1 = g2[k + 32];
                    a0 = fma(x0.x, h0.x, fma(x0.y, h0.y, a0)); a1 = fma(x1.x, h1.x, fma(x1.y, h1.y, a1));
                }
                for (; k < ldw / 2; k += 32) { const double2 x0 = row[k], h0 = g2[k]; a0 = fma(x0.x, h0.x, fma(x0.y, h0.y, a0)); }
                const double sa = warp_sum(a0 + a1);
                if (lane == 0) out[(size_t)s * m + c * CH + r] = sa;
            }
            __syncthreads();
        }
    }
}

__global__ void fill_kernel(double *W, size_t n, int fill) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        const double v = (double)(h & 0xfffff) / 1048576.0 - 0.5;
        W[i] = (fill == 2 || (h >> 40) % 100 < 29) ? v : 0.0;
    }
}
__global__ void fill_g(double *g, int n) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) g[i] = 0.001 * (i % 97) - 0.04; }

int main(int argc, char **argv) {
    const int S = argc > 1 ? atoi(argv[1]) : 2048, m = argc > 2 ? atoi(argv[2]) : 1242;
    const int ldw = (m + 1) & ~1;
    const size_t wcount = (size_t)S * m * ldw;
    double *W, *g, *out; int *queue;
    CK(cudaMalloc(&W, wcount * 8)); CK(cudaMalloc(&g, ldw * 8)); CK(cudaMalloc(&out, (size_t)S * m * 8)); CK(cudaMalloc(&queue, 4));
    CK(cudaMemset(W, 0, wcount * 8)); CK(cudaMemset(g, 0, ldw * 8));
    const int fill = argc > 3 ? atoi(argv[3]) : 0;   // 0: zeros; 1: pseudo-random doubles, 29 % non-zero like a basis inverse; 2: all non-zero
    if (fill) { fill_kernel<<<4096, 256>>>(W, wcount, fill); fill_g<<<8, 256>>>(g, ldw); CK(cudaDeviceSynchronize()); }
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = (double)wcount * 8 / 1e9;
    auto run = [&](const char *name, auto launch) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaMemset(queue, 0, 4));
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%-40s %8.3f ms  %8.1f GB/s\n", name, best, gb / (best / 1e3));
    };
    const size_t sg = (size_t)(ldw + 16) * 8;
    run("v0 scalar 2 rows x 4 segs, 1 CTA/SM", [&] { v0<<<sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    {
        const size_t big = 215 * 1024;
        CK(cudaFuncSetAttribute(v0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));
        run("v0 with 215 KB of shared memory", [&] { v0<<<sms, 512, big>>>(W, g, out, S, m, ldw, queue); });
        CK(cudaFuncSetAttribute(v0p<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));
        CK(cudaFuncSetAttribute(v0p<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));
        CK(cudaFuncSetAttribute(v0p<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));
        CK(cudaFuncSetAttribute(v0p<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));
        CK(cudaFuncSetAttribute(v0p<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big));
        run("v0p default, 215 KB smem", [&] { v0p<0><<<sms, 512, big>>>(W, g, out, S, m, ldw, queue); });
        run("v0p ld.cg, 215 KB smem", [&] { v0p<1><<<sms, 512, big>>>(W, g, out, S, m, ldw, queue); });
        run("v0p ld.cs, 215 KB smem", [&] { v0p<2><<<sms, 512, big>>>(W, g, out, S, m, ldw, queue); });
        run("v0p ld.lu, 215 KB smem", [&] { v0p<3><<<sms, 512, big>>>(W, g, out, S, m, ldw, queue); });
        run("v0p ld.nc.L1::no_allocate, 215 KB smem", [&] { v0p<4><<<sms, 512, big>>>(W, g, out, S, m, ldw, queue); });
        for (int kb : {32, 48, 64, 80, 100, 111, 132, 148, 164, 196}) {
            char name[64]; snprintf(name, sizeof name, "v0p default, %d KB smem", kb);
            const size_t sz = (size_t)kb * 1024;
            run(name, [&] { v0p<0><<<sms, 512, sz>>>(W, g, out, S, m, ldw, queue); });
        }
        run("v0p ld.cg, 10 KB smem", [&] { v0p<1><<<sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    }
    run("v0 same, 2x CTAs (grid 2 per SM)", [&] { v0<<<2 * sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    run("v1 double2 unr4, 1 CTA/SM", [&] { v1<4, 1><<<sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    run("v1 double2 unr8, 1 CTA/SM", [&] { v1<8, 1><<<sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    run("v1 double2 unr4, 2 CTA/SM", [&] { v1<4, 2><<<2 * sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    run("v1 double2 unr8, 2 CTA/SM", [&] { v1<8, 2><<<2 * sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    run("v1 double2 unr4, 4 CTA/SM", [&] { v1<4, 2><<<4 * sms, 512, sg>>>(W, g, out, S, m, ldw, queue); });
    {
        constexpr int ST = 4, CH = 2;
        const size_t sm3 = (size_t)((ldw + 15) & ~15) * 8 + (size_t)ST * CH * ldw * 8 + 128;
        CK(cudaFuncSetAttribute(v3<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        run("v3 bulk copies 4 stages x 2 rows (80 KB)", [&] { v3<ST, CH><<<sms, 512, sm3>>>(W, g, out, S, m, ldw, queue); });
    }
    {
        constexpr int ST = 5, CH = 2;
        const size_t sm3 = (size_t)((ldw + 15) & ~15) * 8 + (size_t)ST * CH * ldw * 8 + 128;
        CK(cudaFuncSetAttribute(v3<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        run("v3 bulk copies 5 stages x 2 rows (100 KB)", [&] { v3<ST, CH><<<sms, 512, sm3>>>(W, g, out, S, m, ldw, queue); });
    }
    {
        constexpr int ST = 8, CH = 2;
        const size_t sm3 = (size_t)((ldw + 15) & ~15) * 8 + (size_t)ST * CH * ldw * 8 + 128;
        CK(cudaFuncSetAttribute(v3<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        run("v3 bulk copies 8 stages x 2 rows (160 KB)", [&] { v3<ST, CH><<<sms, 512, sm3>>>(W, g, out, S, m, ldw, queue); });
    }
    {
        constexpr int ST = 4, CH = 4;
        const size_t sm3 = (size_t)((ldw + 15) & ~15) * 8 + (size_t)ST * CH * ldw * 8 + 128;
        CK(cudaFuncSetAttribute(v3<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        run("v3 bulk copies 4 stages x 4 rows (160 KB)", [&] { v3<ST, CH><<<sms, 512, sm3>>>(W, g, out, S, m, ldw, queue); });
    }
    return 0;
}
