// micro-benchmark: what bounds  acc += (double)v16 * p  on sm_100a - the fp16 -> fp64 conversion (F2F), the DFMA, or neither.
// 148 CTAs x 512 threads (the PV loop's shape).  Variants: 0 = DFMA only; 1 = cvt via F2F.F64.F32; 2 = doubles built with
// integer ops from the fp16 bits (normal numbers only); 3 = as 2 plus the sub-normal / zero fix (exact for every finite fp16).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ double h2d_int(uint32_t h) { // normal fp16 -> double, bits only
    const uint32_t hi = ((h & 0x7fffu) << 10) + 0x3F000000u | ((h & 0x8000u) << 16);
    return __hiloint2double((int)hi, 0);
}
__device__ __forceinline__ double h2d_int_full(uint32_t h) { // + zero / sub-normal: exponent field 0 -> build with field 1, subtract 2^-14
    const uint32_t mag = h & 0x7fffu;
    const bool sub = mag < 0x0400u;
    const uint32_t m2 = sub ? (mag | 0x0400u) : mag;
    const uint32_t hi = (m2 << 10) + 0x3F000000u;
    double d = __hiloint2double((int)hi, 0);
    if (sub) d = __dadd_rn(d, -6.103515625e-05); // 2^-14
    return (h & 0x8000u) ? -d : d;
}
template <int V> __global__ void __launch_bounds__(512) k(const uint4 *in, double *out, long long *cyc, int n) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint4 raw = in[threadIdx.x];
    const double pd = 0.5 + threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            double a, b;
            if (V == 0) { a = __hiloint2double(w[e], 1); b = __hiloint2double(w[e], 2); }
            else if (V == 1) { const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[e])); a = (double)f.x; b = (double)f.y; }
            else if (V == 2) { a = h2d_int(w[e] & 0xffffu); b = h2d_int(w[e] >> 16); }
            else { a = h2d_int_full(w[e] & 0xffffu); b = h2d_int_full(w[e] >> 16); }
            acc[2 * e] = __fma_rn(a, pd, acc[2 * e]);
            acc[2 * e + 1] = __fma_rn(b, pd, acc[2 * e + 1]);
        }
        raw.x += 0x00010001u; raw.y ^= raw.x; raw.z += raw.y & 0x03ff03ffu; raw.w ^= 0x00100010u;
    }
    long long t1 = clock64();
    double s = 0;
    for (int e = 0; e < 8; e++) s += acc[e];
    out[blockIdx.x * 512 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    uint4 *in; double *out; long long *cyc;
    cudaMalloc(&in, 512 * 16); cudaMemset(in, 0x3c, 512 * 16); cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
    const int n = 2000;
    long long h[148];
    for (int v = 0; v < 4; v++) {
        for (int rep = 0; rep < 2; rep++) {
            if (v == 0) k<0><<<148, 512>>>(in, out, cyc, n);
            if (v == 1) k<1><<<148, 512>>>(in, out, cyc, n);
            if (v == 2) k<2><<<148, 512>>>(in, out, cyc, n);
            if (v == 3) k<3><<<148, 512>>>(in, out, cyc, n);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("variant %d: %.1f cycles per iteration (8 values per thread, 512 threads per SM) -> %.1f values/clk/SM  [%s]\n", v,
               (double)h[0] / n, 8.0 * 512 * n / (double)h[0], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
