// dependent-issue latencies on sm_100a: DADD, DMUL, DFMA chains, LDS.128, SHFL, bar.sync with n warps
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dadd(double *o, double a, int n, long long *t) {
    double x = a; long long t0 = clock64();
    #pragma unroll 64
    for (int i = 0; i < n; ++i) x = __dadd_rn(x, a);
    long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) t[0] = t1 - t0;
}
__global__ void k_dmul(double *o, double a, int n, long long *t) {
    double x = a; long long t0 = clock64();
    #pragma unroll 64
    for (int i = 0; i < n; ++i) x = __dmul_rn(x, a);
    long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) t[0] = t1 - t0;
}
__global__ void k_dfma(double *o, double a, int n, long long *t) {
    double x = a; long long t0 = clock64();
    #pragma unroll 64
    for (int i = 0; i < n; ++i) x = __fma_rn(x, a, a);
    long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) t[0] = t1 - t0;
}
// nchain independent DADD chains per thread
template <int C> __global__ void k_dadd_ilp(double *o, double a, int n, long long *t) {
    double x[C]; for (int c = 0; c < C; ++c) x[c] = a + c; long long t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) {
        #pragma unroll
        for (int c = 0; c < C; ++c) x[c] = __dadd_rn(x[c], a);
    }
    long long t1 = clock64(); double s = 0; for (int c = 0; c < C; ++c) s += x[c]; o[threadIdx.x] = s; if (!threadIdx.x) t[0] = t1 - t0;
}
__global__ void k_lds(double *o, int n, long long *t) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
    __syncthreads();
    int p = threadIdx.x; long long t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) p = idx[p];
    long long t1 = clock64(); o[threadIdx.x] = p; if (!threadIdx.x) t[0] = t1 - t0;
}
__global__ void k_shfl(double *o, double a, int n, long long *t) {
    double x = a + threadIdx.x; long long t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < n; ++i) x = __shfl_down_sync(0xffffffffu, x, 1);
    long long t1 = clock64(); o[threadIdx.x] = x; if (!threadIdx.x) t[0] = t1 - t0;
}
__global__ void k_bar(double *o, int n, long long *t) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
    long long t1 = clock64(); o[threadIdx.x] = 0; if (!threadIdx.x) t[0] = t1 - t0;
}
int main() {
    double *o; long long *t, h; cudaMalloc(&o, 8 * 1024); cudaMalloc(&t, 8);
    const int n = 4096;
#define RUN(name, call, thr) for (int r = 0; r < 2; ++r) { call; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost); } printf("%-28s threads %4d: %.2f cycles/op\n", name, thr, (double)h / n);
    RUN("DADD chain", (k_dadd<<<1, 32>>>(o, 1.0000001, n, t)), 32)
    RUN("DMUL chain", (k_dmul<<<1, 32>>>(o, 1.0000001, n, t)), 32)
    RUN("DFMA chain", (k_dfma<<<1, 32>>>(o, 1.0000001, n, t)), 32)
    RUN("DADD 2 chains (per iter)", (k_dadd_ilp<2><<<1, 32>>>(o, 1.0000001, n, t)), 32)
    RUN("DADD 4 chains (per iter)", (k_dadd_ilp<4><<<1, 32>>>(o, 1.0000001, n, t)), 32)
    RUN("DADD 8 chains (per iter)", (k_dadd_ilp<8><<<1, 32>>>(o, 1.0000001, n, t)), 32)
    RUN("DADD chain, 4 warps", (k_dadd<<<1, 128>>>(o, 1.0000001, n, t)), 128)
    RUN("DADD chain, 8 warps", (k_dadd<<<1, 256>>>(o, 1.0000001, n, t)), 256)
    RUN("DADD chain, 16 warps", (k_dadd<<<1, 512>>>(o, 1.0000001, n, t)), 512)
    RUN("DADD chain, 32 warps", (k_dadd<<<1, 1024>>>(o, 1.0000001, n, t)), 1024)
    RUN("DADD 4 chains, 16 warps", (k_dadd_ilp<4><<<1, 512>>>(o, 1.0000001, n, t)), 512)
    RUN("LDS pointer chase", (k_lds<<<1, 32>>>(o, n, t)), 32)
    RUN("SHFL chain (f64 = 2 shfl)", (k_shfl<<<1, 32>>>(o, 1.0, n, t)), 32)
    RUN("bar.sync 1 warp", (k_bar<<<1, 32>>>(o, n, t)), 32)
    RUN("bar.sync 10 warps", (k_bar<<<1, 320>>>(o, n, t)), 320)
    RUN("bar.sync 14 warps", (k_bar<<<1, 448>>>(o, n, t)), 448)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
