#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double *out, long long *cyc, double a, double b, int n) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lat4(double *out, long long *cyc, double a, double b, int n) {
    double x = a + threadIdx.x, y = x + 1, z = x + 2, w = x + 3;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = fma(x, b, a); y = fma(y, b, a); z = fma(z, b, a); w = fma(w, b, a); }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x + y + z + w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rcp(double *out, long long *cyc, double a, int n) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8 * 4096);
    int n = 4096;
    for (int warps : {1, 2, 4, 8}) {
        k_lat<<<1, 32 * warps>>>(out, cyc, 1.0, 0.999, n); cudaDeviceSynchronize();
        k_lat<<<1, 32 * warps>>>(out, cyc, 1.0, 0.999, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent DFMA chain, %d warps/SM (%d per SMSP): %.2f cycles per DFMA per warp\n", warps, (warps + 3) / 4, (double)h / (4.0 * n));
    }
    k_lat4<<<1, 32>>>(out, cyc, 1.0, 0.999, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 independent DFMA chains, 1 warp: %.2f cycles per DFMA\n", (double)h / (4.0 * n));
    k_lat4<<<1, 128 * 4>>>(out, cyc, 1.0, 0.999, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 independent DFMA chains, 16 warps (4/SMSP): %.2f cycles per DFMA per warp\n", (double)h / (4.0 * n));
    k_rcp<<<1, 32>>>(out, cyc, 1.0, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent rcp.approx.f64 + DADD: %.2f cycles per pair\n", (double)h / n);
    return 0;
}
