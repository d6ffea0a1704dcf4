// micro-benchmark: instruction-cache capacity on sm_100a.  A loop whose body is N independent-ish FFMAs
// (16 B each) is run many times; cycles/instruction jumps when the body no longer fits a cache level.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void k(float *out, long long *cyc, int iters) {
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
    const float b = 1.0001f, c = 0.5f;
    long long t0 = 0;
    for (int it = 0; it < iters; it++) {
        if (it == 1) t0 = clock64(); // skip the cold first pass
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int N> void run(float *d, long long *c, int threads) {
    const int iters = 21;
    k<N><<<1, threads>>>(d, c, iters);
    long long h;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("body %6d instr (%5d KB) threads %4d: %.2f cyc/instr/warp  (first pass excluded)\n", N, N * 16 / 1024, threads,
           (double)h / ((iters - 1) * (double)N));
}
int main() {
    float *d; long long *c;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 64);
    for (int threads : {32, 416}) {
        run<512>(d, c, threads); run<1024>(d, c, threads); run<2048>(d, c, threads); run<4096>(d, c, threads);
        run<6144>(d, c, threads); run<8192>(d, c, threads); run<12288>(d, c, threads); run<16384>(d, c, threads);
        run<32768>(d, c, threads);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
