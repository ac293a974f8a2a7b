// Probe: accuracy of the MUFU.RCP64H seed (rcp.approx.ftz.f64) and of the Newton variants built on it.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rcp_probe rcp_probe.cu && ./rcp_probe
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double seed(double a) { double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a)); return x; }
__global__ void probe(double* out, int n) {
    double worst_seed = 0, worst_cubic = 0, worst_two = 0, worst_c1 = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double a = 1.0 + 1023.0 * (double)k / (double)n + 1e-9 * k;   // r^2 range of interest 1..1024
        double exact = 1.0 / a;
        double x = seed(a);
        worst_seed = fmax(worst_seed, fabs(x - exact) / exact);
        double e = fma(-a, x, 1.0);
        double x3 = fma(x, fma(e, e, e), x);                           // cubic step
        worst_cubic = fmax(worst_cubic, fabs(x3 - exact) / exact);
        double e2 = fma(-a, x3, 1.0);
        double x5 = fma(x3, e2, x3);                                   // + quadratic step
        worst_c1 = fmax(worst_c1, fabs(x5 - exact) / exact);
        double y = fma(x, e, x); double f = fma(-a, y, 1.0); y = fma(y, f, y);   // two quadratic steps
        worst_two = fmax(worst_two, fabs(y - exact) / exact);
    }
    atomicMax((unsigned long long*)&out[0], __double_as_longlong(worst_seed));
    atomicMax((unsigned long long*)&out[1], __double_as_longlong(worst_cubic));
    atomicMax((unsigned long long*)&out[2], __double_as_longlong(worst_c1));
    atomicMax((unsigned long long*)&out[3], __double_as_longlong(worst_two));
}
int main() {
    double* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
    probe<<<296, 256>>>(d, 1 << 28);
    double h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("seed %.3e (2^%.1f)  cubic %.3e  cubic+quad %.3e  quad+quad %.3e  (eps=%.3e)\n", h[0], log2(h[0]), h[1], h[2], h[3], 1.11e-16);
    return 0;
}
