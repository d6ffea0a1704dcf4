// micro-benchmark: issue rate of IDP.4A / LOP3 / I2FP / HADD2.F32 / SHFL per SM sub-partition on sm_100a
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_dp4a tools/ubench_dp4a.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(int *out, long long *cyc, int iters, int seed) {
    int a[8], b = seed + threadIdx.x;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (i + 1) + threadIdx.x; f[i] = (float)a[i]; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = __dp4a(a[i], b, a[i]);
                if (OP == 1) a[i] = (a[i] & 0x0F0F0F0F) ^ b;
                if (OP == 2) f[i] = __fmaf_rn(f[i], 1.0001f, f[i]);
                if (OP == 3) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 7));
                if (OP == 4) { f[i] = (float)a[i]; a[i] = __float_as_int(f[i]) + 1; }
                if (OP == 5) { a[i] = __dp4a(a[i] & 0x0F0F0F0F, b, a[i]); }
            }
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + (int)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, int *d, long long *c) {
    for (int threads : {128, 256, 512}) {
        const int iters = 200;
        k<OP><<<1, threads>>>(d, c, iters, 3);
        long long h;
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double per_smsp = (double)iters * 64 * (threads / 128); // warp-instructions per SMSP
        printf("%-28s warps/SMSP %d: %.2f cycles per warp-instruction per SMSP\n", name, threads / 128, (double)h / per_smsp);
    }
}
int main() {
    int *d; long long *c;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 64);
    run<0>("IDP.4A", d, c);
    run<1>("LOP3 x2", d, c);
    run<2>("FFMA", d, c);
    run<3>("SHFL.BFLY", d, c);
    run<4>("I2F + IADD", d, c);
    run<5>("LOP3 + IDP.4A", d, c);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
