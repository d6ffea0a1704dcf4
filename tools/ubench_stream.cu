// micro-benchmark: how fast can a persistent sm_100a kernel stream a Q4_0 matrix (K = 2048) through a
// TMA-fed shared-memory ring and dot it with a register-resident Q8_0 activation row?
// Variants are compared against a pure read kernel (LDG.128 sum) to separate "HBM" from "consumer" limits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -o tools/ubench_stream tools/ubench_stream.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ float lds_half(uint32_t addr) {
    unsigned short h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr int NB = 64;            // blocks per row (K = 2048)
constexpr int ROW_PB = NB * 16;   // payload bytes per row
constexpr int ROW_SB = NB * 2;    // scale bytes per row

struct Act { int a[8]; float d; int bias; };

// one block term, reference arithmetic: ((float)sumi * dx) * dy, sumi = sum (q-8)*a
template <int MODE>
__device__ __forceinline__ float block_term(const uint4 w, float dx, const Act &ab) {
    const uint32_t q[4] = {w.x, w.y, w.z, w.w};
    int sumi;
    if (MODE == 0) { // one chain, shift per word
        sumi = ab.bias;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            sumi = __dp4a((int)(q[i] & 0x0F0F0F0Fu), ab.a[i], sumi);
            sumi = __dp4a((int)((q[i] >> 4) & 0x0F0F0F0Fu), ab.a[4 + i], sumi);
        }
    } else { // two chains; the high nibbles stay in place (x16) and are divided out once
        int lo = ab.bias, hi = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            lo = __dp4a((int)(q[i] & 0x0F0F0F0Fu), ab.a[i], lo);
            hi = dp4a_us(q[i] & 0xF0F0F0F0u, ab.a[4 + i], hi);
        }
        sumi = lo + (hi >> 4);
    }
    return __fmul_rn(__fmul_rn((float)sumi, dx), ab.d);
}

// NW consumer warps, each pass = 2 rows x 64 blocks (lane l: blocks l, l+32), PPW passes per warp per tile
template <int NW, int PPW, int SLOTS, int MODE>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
stream_kernel(const uint8_t *qs, const __half *d, const uint8_t *act, float *y, int n_tiles) {
    constexpr int ROWS = 2 * NW * PPW;
    constexpr int SLOT_BYTES = ROWS * (ROW_PB + ROW_SB);
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t ring = smem_u32(smem);
    const uint32_t bars = ring + SLOTS * SLOT_BYTES; // full[SLOTS], empty[SLOTS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < SLOTS; i++) {
            mbar_init(bars + 8 * i, 1);
            mbar_init(bars + 8 * (SLOTS + i), NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int G = gridDim.x, cta = blockIdx.x;
    if (warp == NW) {
        if (lane == 0) {
            int slot = 0; uint32_t par = 0;
            for (int t = cta; t < n_tiles; t += G) {
                mbar_wait(bars + 8 * (SLOTS + slot), par ^ 1u);
                mbar_expect_tx(bars + 8 * slot, SLOT_BYTES);
                const size_t row0 = (size_t)t * ROWS;
                bulk_g2s(ring + slot * SLOT_BYTES, qs + row0 * ROW_PB, ROWS * ROW_PB, bars + 8 * slot);
                bulk_g2s(ring + slot * SLOT_BYTES + ROWS * ROW_PB, d + row0 * NB, ROWS * ROW_SB, bars + 8 * slot);
                if (++slot == SLOTS) { slot = 0; par ^= 1u; }
            }
        }
        return;
    }
    Act ab[2];
#pragma unroll
    for (int b = 0; b < 2; b++) {
        const int blk = lane + 32 * b;
        const uint4 lo = *reinterpret_cast<const uint4 *>(act + blk * 32);
        const uint4 hi = *reinterpret_cast<const uint4 *>(act + blk * 32 + 16);
        ab[b].a[0] = lo.x; ab[b].a[1] = lo.y; ab[b].a[2] = lo.z; ab[b].a[3] = lo.w;
        ab[b].a[4] = hi.x; ab[b].a[5] = hi.y; ab[b].a[6] = hi.z; ab[b].a[7] = hi.w;
        ab[b].d = __half2float(reinterpret_cast<const __half *>(act + 2048)[blk]);
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s = __dp4a(ab[b].a[i], 0x01010101, s);
        ab[b].bias = -8 * s;
    }
    const int half = lane >> 4;
    int slot = 0; uint32_t par = 0;
    for (int t = cta; t < n_tiles; t += G) {
        const uint32_t base = ring + slot * SLOT_BYTES;
        mbar_wait(bars + 8 * slot, par);
#pragma unroll
        for (int p = 0; p < PPW; p++) {
            const int r0 = 2 * (warp + NW * p);
            const uint32_t pay = base + r0 * ROW_PB + lane * 16;
            const uint32_t scl = base + ROWS * ROW_PB + r0 * ROW_SB + lane * 2;
            uint4 w[4]; float dx[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                w[k] = lds_u4(pay + (k >> 1) * ROW_PB + (k & 1) * 512);
                dx[k] = lds_half(scl + (k >> 1) * ROW_SB + (k & 1) * 64);
            }
            float term[4];
#pragma unroll
            for (int k = 0; k < 4; k++) term[k] = block_term<MODE>(w[k], dx[k], ab[k & 1]);
            const float a0 = __fadd_rn(__fadd_rn(0.0f, term[0]), term[1]);
            const float a1 = __fadd_rn(__fadd_rn(0.0f, term[2]), term[3]);
            const float send = half ? a0 : a1;
            float keep = __fadd_rn(half ? a1 : a0, __shfl_xor_sync(0xffffffffu, send, 16));
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) keep = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, keep, off));
            if ((lane & 15) == 0) y[(size_t)t * ROWS + r0 + half] = keep;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (SLOTS + slot));
        if (++slot == SLOTS) { slot = 0; par ^= 1u; }
    }
}

// pure read: every thread sums LDG.128 words of its CTA's contiguous share
__global__ void __launch_bounds__(512) read_kernel(const uint4 *p, size_t n_vec, uint32_t *out) {
    uint32_t acc = 0;
    const size_t per = n_vec / gridDim.x;
    const uint4 *q = p + per * blockIdx.x;
    for (size_t i = threadIdx.x; i < per; i += 4 * blockDim.x) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i + u * blockDim.x < per)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(q + i + u * blockDim.x));
            else v[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// pure TMA: ring of slots, consumers only arrive (no math) — what the copy engine path sustains per SM
template <int SLOTS, int SLOT_BYTES>
__global__ void __launch_bounds__(64, 1) tma_only_kernel(const uint8_t *p, size_t bytes_per_cta, uint32_t *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t ring = smem_u32(smem);
    const uint32_t bars = ring + SLOTS * SLOT_BYTES;
    if (threadIdx.x == 0) {
        for (int i = 0; i < SLOTS; i++) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 8 * (SLOTS + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t *src = p + bytes_per_cta * blockIdx.x;
    const int n = (int)(bytes_per_cta / SLOT_BYTES);
    if (threadIdx.x == 32) {
        int slot = 0; uint32_t par = 0;
        for (int t = 0; t < n; t++) {
            mbar_wait(bars + 8 * (SLOTS + slot), par ^ 1u);
            mbar_expect_tx(bars + 8 * slot, SLOT_BYTES);
            bulk_g2s(ring + slot * SLOT_BYTES, src + (size_t)t * SLOT_BYTES, SLOT_BYTES, bars + 8 * slot);
            if (++slot == SLOTS) { slot = 0; par ^= 1u; }
        }
    } else if (threadIdx.x == 0) {
        int slot = 0; uint32_t par = 0; uint32_t acc = 0;
        for (int t = 0; t < n; t++) {
            mbar_wait(bars + 8 * slot, par);
            acc += lds_u32(ring + slot * SLOT_BYTES);
            mbar_arrive(bars + 8 * (SLOTS + slot));
            if (++slot == SLOTS) { slot = 0; par ^= 1u; }
        }
        if (acc == 0x12345678u) out[0] = acc;
    }
}

static int g_grid = 148;
template <typename F> float time_ms(F f, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

template <int NW, int PPW, int SLOTS, int MODE>
void run_stream(const uint8_t *qs, const __half *d, const uint8_t *act, float *y, size_t rows, std::vector<float> *keep) {
    constexpr int ROWS = 2 * NW * PPW;
    constexpr int SLOT_BYTES = ROWS * (ROW_PB + ROW_SB);
    const int smem = SLOTS * SLOT_BYTES + 16 * SLOTS;
    const int n_tiles = (int)(rows / ROWS);
    auto fn = stream_kernel<NW, PPW, SLOTS, MODE>;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(y, 0, rows * 4));
    const float ms = time_ms([&] { fn<<<g_grid, NW * 32 + 32, smem>>>(qs, d, act, y, n_tiles); }, 10);
    CK(cudaGetLastError());
    const double bytes = (double)n_tiles * SLOT_BYTES;
    std::vector<float> h((size_t)n_tiles * ROWS);
    CK(cudaMemcpy(h.data(), y, h.size() * 4, cudaMemcpyDeviceToHost));
    int bad = -1;
    if (keep->empty()) *keep = h;
    else for (size_t i = 0; i < h.size() && i < keep->size(); i++) if (h[i] != (*keep)[i]) { bad = (int)i; break; }
    printf("grid %3d stream NW=%2d PPW=%d SLOTS=%d MODE=%d tile=%6d B smem=%6d: %.1f us  %.2f TB/s  %s\n", g_grid, NW, PPW, SLOTS, MODE,
           SLOT_BYTES, smem, ms * 1e3, bytes / ms / 1e9, bad < 0 ? "same" : "DIFF");
    printf("     -> %.1f GB/s per CTA\n", bytes / ms / 1e6 / g_grid);
}

int main() {
    const size_t rows = 148 * 32 * 56; // 265 216 rows x 1152 B = 305 MB
    const size_t pb = rows * ROW_PB, sb = rows * ROW_SB;
    uint8_t *qs; __half *d; uint8_t *act; float *y; uint32_t *out;
    CK(cudaMalloc(&qs, pb + sb)); d = (__half *)(qs + pb);
    CK(cudaMalloc(&act, 2048 + 128)); CK(cudaMalloc(&y, rows * 4)); CK(cudaMalloc(&out, 64));
    std::vector<uint8_t> h(pb + sb);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < pb; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint8_t)(s >> 24); }
    __half *hd = (__half *)(h.data() + pb);
    for (size_t i = 0; i < rows * NB; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; hd[i] = __float2half(0.001f + (float)(s & 1023) * 1e-5f); }
    CK(cudaMemcpy(qs, h.data(), pb + sb, cudaMemcpyHostToDevice));
    std::vector<uint8_t> ha(2048 + 128);
    for (int i = 0; i < 2048; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; ha[i] = (uint8_t)(int8_t)((int)(s % 255) - 127); }
    for (int i = 0; i < 64; i++) ((__half *)(ha.data() + 2048))[i] = __float2half(0.01f + 0.001f * i);
    CK(cudaMemcpy(act, ha.data(), ha.size(), cudaMemcpyHostToDevice));

    for (int g : {148, 296, 592, 1184}) {
        const float ms = time_ms([&] { read_kernel<<<g, 512>>>((const uint4 *)qs, (pb + sb) / 16, out); }, 10);
        printf("read_kernel grid %4d: %.1f us  %.2f TB/s\n", g, ms * 1e3, (double)(pb + sb) / ms / 1e9);
    }
    {
        constexpr int SL = 6, SB = 32768;
        auto fn = tma_only_kernel<SL, SB>;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SL * SB + 16 * SL));
        const size_t per = (pb + sb) / 148 / SB * SB;
        const float ms = time_ms([&] { fn<<<148, 64, SL * SB + 16 * SL>>>(qs, per, out); }, 10);
        printf("tma_only 6 x 32 KB: %.1f us  %.2f TB/s\n", ms * 1e3, (double)per * 148 / ms / 1e9);
    }
    {
        constexpr int SL = 12, SB = 16384;
        auto fn = tma_only_kernel<SL, SB>;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SL * SB + 16 * SL));
        const size_t per = (pb + sb) / 148 / SB * SB;
        const float ms = time_ms([&] { fn<<<148, 64, SL * SB + 16 * SL>>>(qs, per, out); }, 10);
        printf("tma_only 12 x 16 KB: %.1f us  %.2f TB/s\n", ms * 1e3, (double)per * 148 / ms / 1e9);
    }
    {
        constexpr int SL = 24, SB = 8192;
        auto fn = tma_only_kernel<SL, SB>;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SL * SB + 16 * SL));
        const size_t per = (pb + sb) / 148 / SB * SB;
        const float ms = time_ms([&] { fn<<<148, 64, SL * SB + 16 * SL>>>(qs, per, out); }, 10);
        printf("tma_only 24 x 8 KB: %.1f us  %.2f TB/s\n", ms * 1e3, (double)per * 148 / ms / 1e9);
    }
    for (int g : {37, 148}) {
        g_grid = g;
        std::vector<float> keep;
        run_stream<16, 1, 5, 1>(qs, d, act, y, rows, &keep);
        run_stream<16, 2, 3, 1>(qs, d, act, y, rows, &keep);
        run_stream<8, 2, 5, 1>(qs, d, act, y, rows, &keep);
        run_stream<8, 4, 3, 1>(qs, d, act, y, rows, &keep);
        run_stream<12, 2, 4, 1>(qs, d, act, y, rows, &keep);
        run_stream<24, 1, 4, 1>(qs, d, act, y, rows, &keep);
        run_stream<28, 1, 3, 1>(qs, d, act, y, rows, &keep);
        run_stream<20, 1, 4, 1>(qs, d, act, y, rows, &keep);
        run_stream<4, 4, 5, 1>(qs, d, act, y, rows, &keep);
        run_stream<4, 8, 3, 1>(qs, d, act, y, rows, &keep);
    }
    return 0;
}
