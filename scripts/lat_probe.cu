// Dependent-chain latencies on B200 (one warp, clock64): DFMA, DADD, DMUL, SHFL(64-bit), LDS.64, MUFU.RCP64H+3DFMA, DSETP+select
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void probe(double* out, long long* t, double seed) {
    __shared__ double sm[64];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 32] = seed;
    __syncthreads();
    double a = seed + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) a = fma(a, b, c);
    t1 = clock64(); if (threadIdx.x == 0) t[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) a = a + c;
    t1 = clock64(); if (threadIdx.x == 0) t[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) a = a * b;
    t1 = clock64(); if (threadIdx.x == 0) t[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) a = __shfl_xor_sync(0xffffffffu, a, 1);
    t1 = clock64(); if (threadIdx.x == 0) t[3] = t1 - t0;
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) { double v = sm[idx & 63]; idx = (int)v + idx; a += v; }   // address depends on the load
    t1 = clock64(); if (threadIdx.x == 0) t[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { double x; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a)); double e = fma(-a, x, 1.0); a = fma(x, fma(e, e, e), x) + 1.5; }
    t1 = clock64(); if (threadIdx.x == 0) t[5] = t1 - t0;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) a = (a > 0.5) ? a - 0.25 : a + 0.75;
    t1 = clock64(); if (threadIdx.x == 0) t[6] = t1 - t0;
    // 4 independent DFMA chains: issue-limited rate of one warp
    double p = a, q = a + 1, r = a + 2, s = a + 3;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { p = fma(p, b, c); q = fma(q, b, c); r = fma(r, b, c); s = fma(s, b, c); }
    t1 = clock64(); if (threadIdx.x == 0) t[7] = t1 - t0;
    out[threadIdx.x] = a + idx + p + q + r + s;
}
int main() {
    double* d; long long* t; cudaMalloc(&d, 256); cudaMalloc(&t, 64);
    probe<<<1, 32>>>(d, t, 0.0); probe<<<1, 32>>>(d, t, 0.0);
    long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
    const char* nm[8] = {"DFMA dep", "DADD dep", "DMUL dep", "SHFL64 dep", "LDS.64 dep(+cvt+add)", "rcp64(MUFU+3DFMA)+DADD dep", "DSETP+sel+DADD dep", "4 indep DFMA (per 4)"};
    for (int k = 0; k < 8; k++) printf("%-28s %.2f cycles\n", nm[k], (double)h[k] / N);
    return 0;
}
