// Dependent-chain latencies on the target GPU (cycles per op, one warp): informs the per-timestep
// critical-path model in DESIGN.md.   nvcc -arch=sm_100a -O3 --fmad=false latency_probe.cu -o probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void probe(double *out, long long *cyc, double a, double b, int idx0) {
    __shared__ double sm[64];
    __shared__ int si[64];
    for (int i = threadIdx.x; i < 64; i += 32) { sm[i] = 1.0 + i * 1e-9; si[i] = (i * 7 + 1) % 64; }
    __syncwarp();
    double x = a; long long t0, t1; int k = 0;
    // 0: DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, b, a);
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 1: DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + b;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 2: DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * b;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 3: compare+select chain (DSETP + FSEL x2)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = (x > a) ? x - b : x + b;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 4: fmax chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fmax(x - b, a);
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 5: division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; i++) x = a / (x + b);
    t1 = clock64(); cyc[k++] = (t1 - t0) * 8;
    // 6: dependent LDS chain (pointer chasing through ints)
    int p = idx0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = si[p];
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 7: LDS.64 + DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + sm[(p + i) & 63];
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 8: shuffle chain (64-bit)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 9: ballot chain
    unsigned m = 0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) m += __ballot_sync(0xffffffffu, (x + m) > a);
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 10: STS + syncwarp + LDS round trip
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + b; __syncwarp(); }
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 11: integer IMAD chain
    int q = p;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) q = q * 3 + idx0;
    t1 = clock64(); cyc[k++] = t1 - t0;
    // 12: predicated branch (divergent) chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) { if ((threadIdx.x + i + q) & 1) x = x + b; else x = x - a; }
    t1 = clock64(); cyc[k++] = t1 - t0;
    out[threadIdx.x] = x + m + p + q;
}
int main() {
    double *out; long long *cyc, h[16];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
    for (int rep = 0; rep < 2; rep++) probe<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3);
    cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    const char *names[] = {"DFMA", "DADD", "DMUL", "DSETP+select", "DADD+fmax", "DADD+DDIV", "LDS(int) dependent", "LDS.64+DADD",
                           "SHFL64+DADD", "ballot+add", "STS+sync+LDS+DADD+sync", "IMAD", "divergent branch+DADD"};
    for (int k = 0; k < 13; k++) printf("%-28s %7.1f cycles/iter\n", names[k], (double)h[k] / N);
    return 0;
}
