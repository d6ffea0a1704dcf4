// ubench_chain: how long does one chain round take for a lone warp?  16 x LDS.128 + 64 dependent FADDs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__global__ void k(float *out, long long *cyc, int other_warps_spin) {
    __shared__ float T[32 * 68 * 3];
    for (int i = threadIdx.x; i < 32 * 68 * 3; i += blockDim.x) T[i] = 1.0f + i * 1e-6f;
    __syncthreads();
    if (threadIdx.x >= 32) {
        if (other_warps_spin) { // busy warps on the same schedulers
            float a = threadIdx.x;
            for (int i = 0; i < 200000; i++) a = a * 1.0001f + 0.5f;
            out[threadIdx.x] = a;
        }
        return;
    }
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(T) + threadIdx.x * 272;
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < 64; r++) {
        float4 t[16];
#pragma unroll
        for (int u = 0; u < 16; u++) t[u] = lds_f4(base + 16 * u + (r % 3) * 8704);
#pragma unroll
        for (int u = 0; u < 16; u++) {
            acc = __fadd_rn(acc, t[u].x); acc = __fadd_rn(acc, t[u].y); acc = __fadd_rn(acc, t[u].z); acc = __fadd_rn(acc, t[u].w);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 64;
}
int main() {
    float *o; long long *c, h;
    cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
    for (int spin = 0; spin < 2; spin++)
        for (int nw : {1, 5, 20}) {
            k<<<1, 32 * nw>>>(o, c, spin); cudaDeviceSynchronize();
            k<<<1, 32 * nw>>>(o, c, spin); cudaDeviceSynchronize();
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("warps %2d other-warps-busy %d: %lld cycles per round (64 dependent FADD + 16 LDS.128)\n", nw, spin, h);
        }
    return 0;
}
