// micro-benchmark: FP64 add latency / throughput, fp32 IEEE division, shuffles on sm_100a (B200)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dadd_chain(double *out, long long *cyc, int n) {
    double a = threadIdx.x * 1e-3, b = 1.000001;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) a = __dadd_rn(a, b);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dadd_ilp(double *out, long long *cyc, int n) {
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7, b = 1.000001;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        a0 = __dadd_rn(a0, b); a1 = __dadd_rn(a1, b); a2 = __dadd_rn(a2, b); a3 = __dadd_rn(a3, b);
        a4 = __dadd_rn(a4, b); a5 = __dadd_rn(a5, b); a6 = __dadd_rn(a6, b); a7 = __dadd_rn(a7, b);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_fdiv(float *out, long long *cyc, int n) {
    float a = 1.0f + threadIdx.x, b = 1.0001f;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) a = __fdiv_rn(a, b);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl(float *out, long long *cyc, int n) {
    float a = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) a = a + __shfl_xor_sync(0xffffffffu, a, 16);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dp4a(int *out, long long *cyc, int n) {
    int a = threadIdx.x, b = 0x01020304, c = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) c = __dp4a(a, b, c);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *d; float *f; int *ii; long long *c, h[4];
    cudaMalloc(&d, 1 << 20); cudaMalloc(&f, 1 << 20); cudaMalloc(&ii, 1 << 20); cudaMalloc(&c, 64);
    const int n = 1024;
    for (int threads : {32, 256, 1024}) {
        k_dadd_chain<<<1, threads>>>(d, c, n); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("DADD chain   threads=%4d: %.1f cyc/op\n", threads, (double)h[0] / n);
        k_dadd_ilp<<<1, threads>>>(d, c, n); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("DADD ilp8    threads=%4d: %.1f cyc/op (per thread-op), %.2f warp-ops/clk/SM\n", threads, (double)h[0] / (n * 8), (threads / 32.0) * n * 8 / h[0]);
        k_fdiv<<<1, threads>>>(f, c, n); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("fdiv_rn chain threads=%4d: %.1f cyc/op\n", threads, (double)h[0] / n);
        k_shfl<<<1, threads>>>(f, c, n); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("shfl+fadd chain threads=%4d: %.1f cyc/op\n", threads, (double)h[0] / n);
        k_dp4a<<<1, threads>>>(ii, c, n); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("dp4a chain   threads=%4d: %.1f cyc/op\n", threads, (double)h[0] / n);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
