// prefill.cu — the one genuinely dense contraction of the path: mul_mat with col_num = T > 1
// (ref: src/hpc.cpp:216-277 called with ne11 = T; graph nodes at src/gemma_model.cpp:692-696, 493, 446-450,
// 740).  Decode (T = 1) is 2 flop per weight byte and HBM-bound; prefill re-uses every weight T times, so it
// belongs on the 5th-generation tensor cores:
//
//   C[T][N] (fp32) = A[T][K] (bf16 activations) x dequant(W[N][K] Q4_0 | Q8_0)^T       (+ residual)
//
// One CTA computes a 256 x 256 tile of C.  Per 64-wide K block
//   * warps 0-15 (PRODUCERS) copy the A tile (bf16, L2) and DEQUANTISE the W tile (int8/int4 codes x fp16 block
//     scale -> bf16, exact product rounded once) into shared memory in the canonical K-major 128-byte-swizzled
//     UMMA layout, fence the async proxy and arrive on the stage's mbarrier;
//   * one lane of warp 16 (MMA ISSUER) waits for the stage and issues 2 x 4 tcgen05.mma (M = 128, N = 256,
//     K = 16, kind::f16, bf16 x bf16 -> fp32) whose accumulators live in TENSOR MEMORY (2 x 256 columns =
//     all 512), then tcgen05.commit releases the stage;
//   * after the last K block the producer warps become the EPILOGUE: tcgen05.ld 32 lanes x 32 columns at a
//     time, add the residual, store fp32.
// Small-N matrices (o-proj, down-proj: 16 tiles) are split along K so that every SM has a tile; the partial
// tiles go to a workspace and are reduced in split order by a second kernel (deterministic).
//
// Numerics: weights q*d are exact in fp32 (8 x 11 bits) and rounded once to bf16; activations are rounded
// to bf16 instead of being re-quantised to Q8_0 per 32 (the reference's INIT stage), so prefill agrees with
// the reference within a tolerance (tests/test_prefill_gpu.py), not bit for bit — the decode path is the
// bit-exact one.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int BM = 256, BN = 256, BK = 64;
constexpr int kStages = 3;
constexpr int kAStage = BM * BK * 2, kBStage = BN * BK * 2; // 32 KB each
constexpr int kStageBytes = kAStage + kBStage;
constexpr int kProducerWarps = 16, kProducerThreads = 512;
constexpr int kGemmThreads = kProducerThreads + 32;
constexpr int kTmemCols = 512;
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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

// UMMA shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row
// atoms of 1024 B (stride-dimension byte offset), descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);       // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading-dimension byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;               // stride-dimension byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: D = F32, A = B = BF16, both K-major, M = 128, N = 256
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { // arrives on the mbarrier when all prior MMAs are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// up to 3 weight matrices that share the A operand run as ONE launch (Q | K | V, gate | up): a segment = one matrix,
// its first row is a multiple of the tile height, so a tile belongs to exactly one segment
constexpr int kMaxSegs = 3;
struct GemmSeg {
    const uint8_t *qs; // weight payload plane [rows][nb][16|32]
    const __half *d;   // weight scale plane [rows][nb]
    float *c;          // [T][ldc] output of this matrix
    int ldc;
    int n_begin, rows; // first row of the segment in the concatenated N, its row count
};
struct GemmArgs {
    GemmSeg seg[kMaxSegs];
    int n_seg;
    int N, K, nb;      // N = concatenated rows (tile-padded between segments), nb = K / 32
    const __nv_bfloat16 *a; // [T][K]
    int T;
    float *ws;         // K-split workspace [ksplit][T][N] or null (ksplit == 1: tiles go straight to seg.c)
    const float *resid; // [T][seg[0].ldc] or null (single segment, ksplit == 1)
    int kblocks_per_split; // 64-wide K blocks per split
};

// int8 / int4 code -> fp32 without the conversion pipe: the (biased) code is dropped into the mantissa of
// 2^23 with one PRMT, the bias is removed with one exact FADD, and one FMUL applies the block scale
// (code * d is exact in fp32: 8 x 11 significant bits); the product is then rounded once to bf16.
template <int SEL> __device__ __forceinline__ float code_to_f32(uint32_t biased_word, float bias, float d) {
    return __fmul_rn(__fadd_rn(__uint_as_float(__byte_perm(biased_word, 0x4B000000u, SEL)), bias), d);
}
__device__ __forceinline__ void word_to_bf16x4(uint32_t biased_word, float bias, float d, uint32_t &o0, uint32_t &o1) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(code_to_f32<0x7540>(biased_word, bias, d), code_to_f32<0x7541>(biased_word, bias, d));
    __nv_bfloat162 p1 = __floats2bfloat162_rn(code_to_f32<0x7542>(biased_word, bias, d), code_to_f32<0x7543>(biased_word, bias, d));
    o0 = *reinterpret_cast<uint32_t *>(&p0);
    o1 = *reinterpret_cast<uint32_t *>(&p1);
}
// 16 bf16 of dequantised weights (two 16-byte chunks of the swizzled row) from 16 int8 codes
__device__ __forceinline__ void dequant16_q8(const uint4 w, float d, uint4 &lo, uint4 &hi) {
    const float bias = -8388736.0f; // -(2^23 + 128): code ^ 0x80 = code + 128
    uint32_t out[8];
    word_to_bf16x4(w.x ^ 0x80808080u, bias, d, out[0], out[1]);
    word_to_bf16x4(w.y ^ 0x80808080u, bias, d, out[2], out[3]);
    word_to_bf16x4(w.z ^ 0x80808080u, bias, d, out[4], out[5]);
    word_to_bf16x4(w.w ^ 0x80808080u, bias, d, out[6], out[7]);
    lo = make_uint4(out[0], out[1], out[2], out[3]);
    hi = make_uint4(out[4], out[5], out[6], out[7]);
}
// one Q4_0 block (16 payload bytes -> 32 weights): elements 0..15 are the low nibbles, 16..31 the high ones
__device__ __forceinline__ void dequant32_q4(const uint4 w, float d, uint4 (&o)[4]) {
    const float bias = -8388616.0f; // -(2^23 + 8): value = nibble - 8
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        word_to_bf16x4(ws[i] & 0x0F0F0F0Fu, bias, d, lo[2 * i], lo[2 * i + 1]);
        word_to_bf16x4((ws[i] >> 4) & 0x0F0F0F0Fu, bias, d, hi[2 * i], hi[2 * i + 1]);
    }
    o[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    o[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    o[2] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    o[3] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
}

__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int TYPE>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_bf16w_kernel(const GemmArgs A) {
    extern __shared__ uint8_t gemm_smem_raw[];
    const uint32_t smem0 = (smem_u32(gemm_smem_raw) + 1023u) & ~1023u; // SWIZZLE_128B atoms need 1024-byte alignment
    const uint32_t bars = smem0 + kStages * kStageBytes;                // full[kStages], empty[kStages], tmem_full
    const uint32_t tmem_slot = bars + 8 * (2 * kStages + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, split = blockIdx.z;
    int si = 0;
    if (A.n_seg > 1 && n0 >= A.seg[1].n_begin) si = (A.n_seg > 2 && n0 >= A.seg[2].n_begin) ? 2 : 1;
    const GemmSeg &sg = A.seg[si];
    const int sn0 = n0 - sg.n_begin; // first row of this tile inside its matrix
    const int kb0 = split * A.kblocks_per_split;
    const int kb1 = min(kb0 + A.kblocks_per_split, (A.K + BK - 1) / BK);

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(bars + 8 * s, kProducerWarps);        // full: one arrive per producer warp
            mbar_init(bars + 8 * (kStages + s), 1);          // empty: tcgen05.commit
        }
        mbar_init(bars + 8 * 2 * kStages, 1);                // accumulators complete
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarps) { // the MMA warp owns tensor memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < kProducerWarps) {
        // ================= producers: thread t owns weight row n0 + t of the tile (64 weights per K block) and
        // 8 of the 2048 16-byte chunks of the A tile.  The raw codes of the NEXT K block are requested before the
        // current one is converted (register double buffer), and the A tile travels with cp.async, so the
        // memory latency of both operands overlaps the dequantisation arithmetic.
        // thread t owns ONE quant block (32 weights) per K block: row n0 + (t >> 1), block 2 kb + (t & 1)
        const int wr = tid >> 1, whalf = tid & 1;
        const int wrow = sn0 + wr;
        const bool wlive = wrow < sg.rows;
        constexpr int PBW = TYPE == GGB_TYPE_Q8_0 ? 2 : 1; // 16-byte payload words per quant block
        // register ring of raw codes, kStages K blocks deep: HBM latency under load (~2000 cycles) is longer
        // than one K block of MMA time, so the codes are requested kStages blocks before they are converted
        uint4 raw[kStages][PBW];
        __half dd[kStages];
        auto fetch = [&](int kb, uint4 (&r)[PBW], __half &d1) {
            if (!wlive || kb >= kb1) return;
            const size_t blk = (size_t)wrow * A.nb + (size_t)kb * 2 + whalf;
            d1 = sg.d[blk];
            const uint4 *src = reinterpret_cast<const uint4 *>(sg.qs + blk * (TYPE == GGB_TYPE_Q8_0 ? 32 : 16));
#pragma unroll
            for (int j = 0; j < PBW; j++) r[j] = __ldg(src + j);
        };
#pragma unroll
        for (int u = 0; u < kStages; u++) {
            dd[u] = __float2half(0.f);
#pragma unroll
            for (int j = 0; j < PBW; j++) raw[u][j] = make_uint4(0, 0, 0, 0);
            fetch(kb0 + u, raw[u], dd[u]);
        }
        uint32_t parity = 0;
        for (int kbase = kb0; kbase < kb1; kbase += kStages, parity ^= 1u) {
#pragma unroll
            for (int u = 0; u < kStages; u++) { // stage == u: the K loop is unrolled by the ring depth
                const int kb = kbase + u;
                if (kb >= kb1) break;
                const int stage = u;
                mbar_wait(bars + 8 * (kStages + stage), parity ^ 1u);
                const uint32_t sa = smem0 + stage * kStageBytes, sb = sa + kAStage;
                // A tile: 256 rows x 128 B; 8 threads per row (16-byte chunks), 64 rows per sweep
                {
                    const int ch = tid & 7;
#pragma unroll
                    for (int r = tid >> 3; r < BM; r += kProducerThreads / 8) {
                        const int row = m0 + r;
                        const uint32_t dst = sa + r * 128 + ((ch ^ (r & 7)) << 4);
                        if (row < A.T) {
                            const void *src = reinterpret_cast<const uint4 *>(A.a + (size_t)row * A.K + (size_t)kb * BK) + ch;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                        } else {
                            sts_u4(dst, make_uint4(0, 0, 0, 0));
                        }
                    }
                }
                // W tile: convert the block requested kStages K blocks ago, then request block kb + kStages
                {
                    uint4 cur[PBW];
#pragma unroll
                    for (int j = 0; j < PBW; j++) cur[j] = raw[u][j];
                    const float d0 = __half2float(dd[u]);
                    fetch(kb + kStages, raw[u], dd[u]);
                    const uint32_t dst = sb + wr * 128;
                    const int sw = wr & 7, c0 = 4 * whalf; // this block's four 16-byte chunks of the 128-byte row
                    uint4 o[4];
                    if (wlive) {
                        if (TYPE == GGB_TYPE_Q8_0) {
                            dequant16_q8(cur[0], d0, o[0], o[1]);
                            dequant16_q8(cur[PBW - 1], d0, o[2], o[3]);
                        } else {
                            dequant32_q4(cur[0], d0, o);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; q++) o[q] = make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int q = 0; q < 4; q++) sts_u4(dst + (((c0 + q) ^ sw) << 4), o[q]);
                }
                // The A copies of THIS block stay in flight (one L2 round trip) while the next block is converted:
                // wait only for the previous block's group, then hand the PREVIOUS stage over.  generic-proxy /
                // cp.async writes -> visible to the tensor core (async proxy) -> arrive.
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");
                if (kb > kb0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + 8 * (stage == 0 ? kStages - 1 : stage - 1));
                }
            }
        }
        if (kb0 < kb1) { // last stage
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * ((kb1 - kb0 - 1) % kStages));
        }
        // ================= epilogue: warp w reads TMEM lanes 32w .. 32w+31 (accumulator rows)
        mbar_wait(bars + 8 * 2 * kStages, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // straight to the matrix's output, or (K-split) into this split's [T][N] plane of the workspace
        float *cbase = A.ws ? A.ws + (size_t)split * A.T * A.N + n0 : sg.c + sn0;
        const int cld = A.ws ? A.N : sg.ldc;
#pragma unroll 1
        for (int mt = 0; mt < BM / 128; mt++) {
            const int lq = warp & 3, chalf = warp >> 2; // TMEM lane quarter of this warp, column quarter it drains
            const int row = m0 + mt * 128 + lq * 32 + lane;
#pragma unroll 1
            for (int c0 = chalf * (BN / 4); c0 < (chalf + 1) * (BN / 4); c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(mt * BN + c0), v);
                if (row < A.T && sn0 + c0 < sg.rows) { // rows % 32 == 0
                    float *dst = cbase + (size_t)row * cld + c0;
                    const float *rs = A.resid ? A.resid + (size_t)row * sg.ldc + sn0 + c0 : nullptr;
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        float4 o = make_float4(__uint_as_float(v[u]), __uint_as_float(v[u + 1]), __uint_as_float(v[u + 2]),
                                               __uint_as_float(v[u + 3]));
                        if (rs) {
                            const float4 r4 = *reinterpret_cast<const float4 *>(rs + u);
                            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                        }
                        *reinterpret_cast<float4 *>(dst + u) = o;
                    }
                }
            }
        }
    } else {
        // ================= MMA issuer (one lane)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, BN);
            int stage = 0;
            uint32_t parity = 0;
            for (int kb = kb0; kb < kb1; kb++) {
                mbar_wait(bars + 8 * stage, parity);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem0 + stage * kStageBytes, sb = sa + kAStage;
                const uint64_t bdesc = umma_desc_sw128(sb);
#pragma unroll
                for (int mt = 0; mt < BM / 128; mt++) {
                    const uint64_t adesc = umma_desc_sw128(sa + mt * 128 * 128);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) // 16 bf16 = 32 B = 2 descriptor units along K
                        umma_bf16(tmem_base + (uint32_t)(mt * BN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                  (kb > kb0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bars + 8 * (kStages + stage)); // frees the stage when these MMAs have read it
                if (++stage == kStages) {
                    stage = 0;
                    parity ^= 1u;
                }
            }
            umma_commit(bars + 8 * 2 * kStages); // accumulators final
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kProducerWarps) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// partial tiles of a K-split GEMM summed in split order (+ residual)
__global__ void splitk_reduce_kernel(const GemmArgs A, int splits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t T = A.T, N = A.N;
    if (i >= T * N) return;
    const int64_t t = i / N;
    const int n = (int)(i % N);
    int si = 0;
    if (A.n_seg > 1 && n >= A.seg[1].n_begin) si = (A.n_seg > 2 && n >= A.seg[2].n_begin) ? 2 : 1;
    const GemmSeg &sg = A.seg[si];
    const int col = n - sg.n_begin;
    if (col >= sg.rows) return; // tile padding between segments
    float acc = A.ws[i];
    for (int s = 1; s < splits; s++) acc += A.ws[(int64_t)s * T * N + i];
    if (A.resid) acc += A.resid[t * sg.ldc + col];
    sg.c[t * sg.ldc + col] = acc;
}

// rms_norm * w -> bf16 (the A operand of the next GEMM): same double-sum order as rms_norm_kernel (elementwise.cu)
__global__ void __launch_bounds__(GGB_DEV_NORM_THREADS)
rms_norm_bf16_kernel(const float *__restrict__ x, const float *__restrict__ w, float eps, int64_t E, __nv_bfloat16 *__restrict__ y) {
    __shared__ double s_w[GGB_DEV_NORM_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *xr = x + (int64_t)blockIdx.x * E;
    double acc = 0.0;
    for (int64_t i = tid; i < E; i += GGB_DEV_NORM_THREADS) {
        const float v = xr[i];
        acc = __dadd_rn(acc, (double)__fmul_rn(v, v));
    }
    acc = warp_sum_f64(acc);
    if (lane == 0) s_w[warp] = acc;
    __syncthreads();
    double sum = s_w[0];
#pragma unroll
    for (int i = 1; i < GGB_DEV_NORM_THREADS / 32; i++) sum = __dadd_rn(sum, s_w[i]);
    const float mean = __double2float_rn(__ddiv_rn(sum, (double)E));
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    for (int64_t i = tid; i < E; i += GGB_DEV_NORM_THREADS)
        y[(int64_t)blockIdx.x * E + i] = __float2bfloat16_rn(__fmul_rn(__fmul_rn(xr[i], scale), w[i]));
}
// gelu(gate) * up -> bf16 (fp16 gelu table like the reference's ggml_gelu; ref: src/gemma_model.cpp:448-449)
__global__ void geglu_bf16_kernel(const float *gate, const float *up, int64_t n, const uint16_t *__restrict__ table,
                                  __nv_bfloat16 *y) {
    // four elements per thread: 128-bit loads of gate / up, four table lookups in flight, one 64-bit store
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 g = *reinterpret_cast<const float4 *>(gate + i), u = *reinterpret_cast<const float4 *>(up + i);
        const uint16_t t0 = table[float_to_half_bits(g.x)], t1 = table[float_to_half_bits(g.y)];
        const uint16_t t2 = table[float_to_half_bits(g.z)], t3 = table[float_to_half_bits(g.w)];
        __nv_bfloat162 lo = __floats2bfloat162_rn(__fmul_rn(half_bits_to_float(t0), u.x), __fmul_rn(half_bits_to_float(t1), u.y));
        __nv_bfloat162 hi = __floats2bfloat162_rn(__fmul_rn(half_bits_to_float(t2), u.z), __fmul_rn(half_bits_to_float(t3), u.w));
        uint2 o;
        o.x = *reinterpret_cast<uint32_t *>(&lo);
        o.y = *reinterpret_cast<uint32_t *>(&hi);
        *reinterpret_cast<uint2 *>(y + i) = o;
    } else {
        for (int64_t j = i; j < n; j++) y[j] = __float2bfloat16_rn(__fmul_rn(half_bits_to_float(table[float_to_half_bits(gate[j])]), up[j]));
    }
}

__global__ void cvt_f32_bf16_kernel(const float *x, int64_t n, __nv_bfloat16 *y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2bfloat16_rn(x[i]);
}

// Causal attention of T query tokens at positions pos0 .. pos0+T-1 over the fp16 KV cache (position-major
// [n_ctx][Hkv*d]).  One warp per (token, GH query heads that share a KV head), the reference's node order
// (ref: src/gemma_model.cpp:454-489): kq = K . f16(q) (fp32 accumulate), soft_max_ext with the fp16 exp table and
// a double sum, P rounded to fp16, kqv = V^T . P.  q is already roped and scaled.  A K / V row is loaded once for
// the GH heads, and the 8 lane partials of a step (8 / GH positions x GH heads) are reduced with ONE folded
// butterfly (9 shuffles instead of 40; the same additions as the plain xor-butterfly).  CUDA cores.
constexpr int kAttnWarps = 4;
template <int GH>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_prefill_kernel(const float *__restrict__ q, const __half *__restrict__ kc, const __half *__restrict__ vc, int T,
                    int pos0, int H, int Hkv, int d, const uint16_t *__restrict__ exp_table, float *__restrict__ out,
                    __nv_bfloat16 *__restrict__ out_bf16, int n_kv_max) {
    constexpr int NP = 8 / GH; // positions per step
    extern __shared__ float attn_s[]; // [kAttnWarps][GH][n_kv_max]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HG = H / GH; // head groups per token
    const int item = blockIdx.x * kAttnWarps + warp;
    if (item >= T * HG) return;
    const int t = item / HG, h0 = (item % HG) * GH, g = h0 / (H / Hkv);
    const int n_kv = pos0 + t + 1;
    float *s = attn_s + (size_t)warp * GH * n_kv_max;
    const size_t stride = (size_t)Hkv * d;
    const bool in = 8 * lane < d;
    float qf[GH][8];
#pragma unroll
    for (int u = 0; u < GH; u++)
#pragma unroll
        for (int e = 0; e < 8; e++)
            qf[u][e] = in ? __half2float(__float2half_rn(q[((size_t)t * H + h0 + u) * d + 8 * lane + e])) : 0.f;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    // after the fold a lane holds value index 4 b4 + 2 b3 + b2 of the step: value = position * GH + head
    const int vi = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
    const int my_p = vi / GH, my_u = vi % GH;
    float mx = -INFINITY; // of head my_u over the positions this lane sees
    for (int j0 = 0; j0 < n_kv; j0 += NP) {
        float v[8];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (in && j0 + p < n_kv) raw = *reinterpret_cast<const uint4 *>(kc + (size_t)(j0 + p) * stride + (size_t)g * d + 8 * lane);
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
            float kf[8];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __half22float2(h2[e]);
                kf[2 * e] = f.x;
                kf[2 * e + 1] = f.y;
            }
#pragma unroll
            for (int u = 0; u < GH; u++) {
                float acc = 0.f;
#pragma unroll
                for (int e = 0; e < 8; e++) acc = fmaf(kf[e], qf[u][e], acc);
                v[p * GH + u] = acc;
            }
        }
        float k4[4], k2[2];
#pragma unroll
        for (int i = 0; i < 4; i++) k4[i] = (b4 ? v[4 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, b4 ? v[i] : v[4 + i], 16);
#pragma unroll
        for (int i = 0; i < 2; i++) k2[i] = (b3 ? k4[2 + i] : k4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? k4[i] : k4[2 + i], 8);
        float k1 = (b2 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? k2[0] : k2[1], 4);
        k1 += __shfl_xor_sync(0xffffffffu, k1, 2);
        k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
        if (j0 + my_p < n_kv) {
            if ((lane & 3) == 0) s[(size_t)my_u * n_kv_max + j0 + my_p] = k1;
            mx = fmaxf(mx, k1);
        }
    }
    // the maximum of every head: lanes with the same head index hold its positions' maxima
    float mxu[GH];
#pragma unroll
    for (int u = 0; u < GH; u++) mxu[u] = warp_max_f32(my_u == u ? mx : -INFINITY);
    __syncwarp();
    float inv[GH];
#pragma unroll
    for (int u = 0; u < GH; u++) {
        float *su = s + (size_t)u * n_kv_max;
        double sum = 0.0;
        for (int j = lane; j < n_kv; j += 32) {
            const float e = half_bits_to_float(exp_table[float_to_half_bits(su[j] - mxu[u])]);
            su[j] = e;
            sum += (double)e;
        }
        sum = warp_sum_f64(sum);
        inv[u] = (float)(1.0 / sum);
    }
    __syncwarp();
    float o[GH][8];
#pragma unroll
    for (int u = 0; u < GH; u++)
#pragma unroll
        for (int e = 0; e < 8; e++) o[u][e] = 0.f;
    for (int j = 0; j < n_kv; j++) { // (unrolling this loop by 8 measured slower: 3.1 vs 2.5 ms per prefill)
        float vf[8];
        if (in) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(vc + (size_t)j * stride + (size_t)g * d + 8 * lane);
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __half22float2(h2[e]);
                vf[2 * e] = f.x;
                vf[2 * e + 1] = f.y;
            }
        }
#pragma unroll
        for (int u = 0; u < GH; u++) {
            const float p = __half2float(__float2half_rn(s[(size_t)u * n_kv_max + j] * inv[u])); // ggml converts P to fp16 for the V mat-mul
            if (in) {
#pragma unroll
                for (int e = 0; e < 8; e++) o[u][e] = fmaf(vf[e], p, o[u][e]);
            }
        }
    }
    if (in) {
#pragma unroll
        for (int u = 0; u < GH; u++) {
            const size_t at = ((size_t)t * H + h0 + u) * d + 8 * lane;
            if (out) {
#pragma unroll
                for (int e = 0; e < 8; e++) out[at + e] = o[u][e];
            }
            if (out_bf16) { // the o-projection's A operand, rounded exactly as ggb_op_cvt_f32_bf16 would
                uint4 pk;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(o[u][0], o[u][1]), p1 = __floats2bfloat162_rn(o[u][2], o[u][3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(o[u][4], o[u][5]), p3 = __floats2bfloat162_rn(o[u][6], o[u][7]);
                pk.x = *reinterpret_cast<uint32_t *>(&p0); pk.y = *reinterpret_cast<uint32_t *>(&p1);
                pk.z = *reinterpret_cast<uint32_t *>(&p2); pk.w = *reinterpret_cast<uint32_t *>(&p3);
                *reinterpret_cast<uint4 *>(out_bf16 + at) = pk;
            }
        }
    }
}

} // namespace

// The MQA / GQA form: ONE CTA (4 warps) per (token, 4 query heads of one KV head).  The attention above is bound by
// L2 bandwidth (every warp re-reads the K / V rows: 1 GB per layer at T = 512); here a K / V row is loaded once for the
// 4 heads and the positions are dealt to the 4 warps, so the traffic drops 4x at the same number of warps.  Scores:
// 2 positions x 4 heads per folded butterfly; max / sum / P.V partials are combined across the warps in shared memory.
__global__ void __launch_bounds__(128)
attn_prefill_gqa_kernel(const float *__restrict__ q, const __half *__restrict__ kc, const __half *__restrict__ vc, int T,
                        int pos0, int H, int Hkv, int d, const uint16_t *__restrict__ exp_table, float *__restrict__ out,
                        __nv_bfloat16 *__restrict__ out_bf16, int n_kv_max) {
    constexpr int GH = 4, NW = 4;
    extern __shared__ float attn_s[]; // s[GH][n_kv_max] | red[NW][GH][256] | wmax[NW][GH] | wsum double[NW][GH]
    float *s = attn_s;
    float *red = attn_s + (size_t)GH * n_kv_max;
    float *wmax = red + NW * GH * 256;
    double *wsum = reinterpret_cast<double *>(wmax + NW * GH);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int HG = H / GH;
    const int t = blockIdx.x / HG, h0 = (blockIdx.x % HG) * GH, g = h0 / (H / Hkv);
    const int n_kv = pos0 + t + 1;
    const size_t stride = (size_t)Hkv * d;
    const bool in = 8 * lane < d;
    float qf[GH][8];
#pragma unroll
    for (int u = 0; u < GH; u++)
#pragma unroll
        for (int e = 0; e < 8; e++)
            qf[u][e] = in ? __half2float(__float2half_rn(q[((size_t)t * H + h0 + u) * d + 8 * lane + e])) : 0.f;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    const int my_p = b4 ? 1 : 0, my_u = (b3 ? 2 : 0) + (b2 ? 1 : 0); // value index = position * GH + head
    float mx = -INFINITY;
    for (int j0 = 2 * warp; j0 < n_kv; j0 += 2 * NW) {
        float v[8];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (in && j0 + p < n_kv) raw = *reinterpret_cast<const uint4 *>(kc + (size_t)(j0 + p) * stride + (size_t)g * d + 8 * lane);
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
            float kf[8];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __half22float2(h2[e]);
                kf[2 * e] = f.x;
                kf[2 * e + 1] = f.y;
            }
#pragma unroll
            for (int u = 0; u < GH; u++) {
                float acc = 0.f;
#pragma unroll
                for (int e = 0; e < 8; e++) acc = fmaf(kf[e], qf[u][e], acc);
                v[p * GH + u] = acc;
            }
        }
        float k4[4], k2[2];
#pragma unroll
        for (int i = 0; i < 4; i++) k4[i] = (b4 ? v[4 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, b4 ? v[i] : v[4 + i], 16);
#pragma unroll
        for (int i = 0; i < 2; i++) k2[i] = (b3 ? k4[2 + i] : k4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? k4[i] : k4[2 + i], 8);
        float k1 = (b2 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? k2[0] : k2[1], 4);
        k1 += __shfl_xor_sync(0xffffffffu, k1, 2);
        k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
        if (j0 + my_p < n_kv) {
            if ((lane & 3) == 0) s[(size_t)my_u * n_kv_max + j0 + my_p] = k1;
            mx = fmaxf(mx, k1);
        }
    }
#pragma unroll
    for (int u = 0; u < GH; u++) {
        const float m = warp_max_f32(my_u == u ? mx : -INFINITY);
        if (lane == 0) wmax[warp * GH + u] = m;
    }
    __syncthreads();
    float inv[GH];
#pragma unroll
    for (int u = 0; u < GH; u++) {
        const float m = fmaxf(fmaxf(wmax[u], wmax[GH + u]), fmaxf(wmax[2 * GH + u], wmax[3 * GH + u]));
        float *su = s + (size_t)u * n_kv_max;
        double sum = 0.0;
        for (int j = tid; j < n_kv; j += 128) {
            const float e = half_bits_to_float(exp_table[float_to_half_bits(su[j] - m)]);
            su[j] = e;
            sum += (double)e;
        }
        sum = warp_sum_f64(sum);
        if (lane == 0) wsum[warp * GH + u] = sum;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < GH; u++) inv[u] = (float)(1.0 / (wsum[u] + wsum[GH + u] + wsum[2 * GH + u] + wsum[3 * GH + u]));
    float o[GH][8];
#pragma unroll
    for (int u = 0; u < GH; u++)
#pragma unroll
        for (int e = 0; e < 8; e++) o[u][e] = 0.f;
    for (int j = warp; j < n_kv; j += NW) {
        float vf[8];
        if (in) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(vc + (size_t)j * stride + (size_t)g * d + 8 * lane);
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __half22float2(h2[e]);
                vf[2 * e] = f.x;
                vf[2 * e + 1] = f.y;
            }
        }
#pragma unroll
        for (int u = 0; u < GH; u++) {
            const float p = __half2float(__float2half_rn(s[(size_t)u * n_kv_max + j] * inv[u])); // P as fp16, like ggml
            if (in) {
#pragma unroll
                for (int e = 0; e < 8; e++) o[u][e] = fmaf(vf[e], p, o[u][e]);
            }
        }
    }
    if (in) {
#pragma unroll
        for (int u = 0; u < GH; u++)
#pragma unroll
            for (int e = 0; e < 8; e++) red[(warp * GH + u) * 256 + 8 * lane + e] = o[u][e];
    }
    __syncthreads();
    if (in) { // warp w finishes head h0 + w: the four warps' partial rows in warp order
        float r[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int c = 8 * lane + e;
            r[e] = ((red[(0 * GH + warp) * 256 + c] + red[(1 * GH + warp) * 256 + c]) + red[(2 * GH + warp) * 256 + c]) +
                   red[(3 * GH + warp) * 256 + c];
        }
        const size_t at = ((size_t)t * H + h0 + warp) * d + 8 * lane;
        if (out) {
#pragma unroll
            for (int e = 0; e < 8; e++) out[at + e] = r[e];
        }
        if (out_bf16) {
            uint4 pk;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(r[0], r[1]), p1 = __floats2bfloat162_rn(r[2], r[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(r[4], r[5]), p3 = __floats2bfloat162_rn(r[6], r[7]);
            pk.x = *reinterpret_cast<uint32_t *>(&p0); pk.y = *reinterpret_cast<uint32_t *>(&p1);
            pk.z = *reinterpret_cast<uint32_t *>(&p2); pk.w = *reinterpret_cast<uint32_t *>(&p3);
            *reinterpret_cast<uint4 *>(out_bf16 + at) = pk;
        }
    }
}

extern "C" {

size_t ggb_gemm_workspace_bytes(int64_t T, int64_t N, int64_t K) {
    (void)K;
    return (size_t)8 * (size_t)T * (size_t)N * 4; // up to 8 K-splits
}

int ggb_op_cvt_f32_bf16(const float *x, int64_t n, void *y_bf16) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && y_bf16 && n > 0, "bad cvt args");
    cvt_f32_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ggb_cuda_stream()>>>(x, n, (__nv_bfloat16 *)y_bf16);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_rms_norm_bf16(const float *x, const float *w, float eps, int64_t E, int64_t T, void *y_bf16) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && w && y_bf16 && E > 0 && T > 0, "bad rms_norm_bf16 args");
    rms_norm_bf16_kernel<<<(unsigned)T, GGB_DEV_NORM_THREADS, 0, ggb_cuda_stream()>>>(x, w, eps, E, (__nv_bfloat16 *)y_bf16);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_geglu_bf16(const float *gate, const float *up, int64_t n, void *y_bf16) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(gate && up && y_bf16 && n > 0, "bad geglu_bf16 args");
    geglu_bf16_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, ggb_cuda_stream()>>>(gate, up, n, ggb_gelu_table_dev(), (__nv_bfloat16 *)y_bf16);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

static int gemm_launch(int n, const ggb_weight *const *w, const void *a_bf16, int64_t T, float *const *c, const int64_t *ldc,
                       const float *resid, void *workspace, size_t workspace_bytes) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(n >= 1 && n <= kMaxSegs && w && a_bf16 && c && ldc && T > 0, "bad gemm args");
    GGB_CHECK_ARG(n == 1 || !resid, "gemm: a residual needs a single matrix");
    const int64_t K = w[0]->cols;
    const int type = w[0]->type;
    GGB_CHECK_ARG(type == GGB_TYPE_Q4_0 || type == GGB_TYPE_Q8_0, "gemm: weight type %d", type);
    GemmArgs A{};
    int64_t N = 0;
    for (int i = 0; i < n; i++) {
        GGB_CHECK_ARG(w[i] && c[i] && w[i]->cols == K && w[i]->type == type, "gemm: matrices of one launch share K and the type");
        const int64_t rows = w[i]->rows;
        GGB_CHECK_ARG(K % 64 == 0 && rows % 32 == 0 && ldc[i] >= rows && ldc[i] % 4 == 0,
                      "gemm needs K %% 64 == 0, N %% 32 == 0 (N=%lld K=%lld)", (long long)rows, (long long)K);
        A.seg[i] = GemmSeg{w[i]->qs, w[i]->d, c[i], (int)ldc[i], (int)N, (int)rows};
        N += i + 1 < n ? (rows + BN - 1) / BN * BN : rows; // the next matrix starts on a tile boundary
    }
    A.n_seg = n;
    static bool attr_done = false;
    if (!attr_done) {
        GGB_CUDA(cudaFuncSetAttribute(gemm_bf16w_kernel<GGB_TYPE_Q4_0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        GGB_CUDA(cudaFuncSetAttribute(gemm_bf16w_kernel<GGB_TYPE_Q8_0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        attr_done = true;
    }
    const int n_tiles = (int)((N + BN - 1) / BN), m_tiles = (int)((T + BM - 1) / BM), kblocks = (int)(K / BK);
    // split K until every SM has a tile (partials are reduced in split order: deterministic)
    int splits = 1;
    const int sms = ggb_sm_count();
    while (splits < 8 && n_tiles * m_tiles * splits * 2 <= sms && kblocks / (splits * 2) >= 4) splits *= 2;
    if (splits > 1 && (!workspace || workspace_bytes < (size_t)splits * T * N * 4)) splits = 1;
    A.N = (int)N; A.K = (int)K; A.nb = (int)w[0]->nb;
    A.a = (const __nv_bfloat16 *)a_bf16; A.T = (int)T;
    A.kblocks_per_split = (kblocks + splits - 1) / splits;
    A.ws = splits > 1 ? (float *)workspace : nullptr;
    A.resid = resid;
    const dim3 grid((unsigned)n_tiles, (unsigned)m_tiles, (unsigned)splits);
    if (splits > 1) A.resid = nullptr; // added by the reducer
    if (type == GGB_TYPE_Q4_0) gemm_bf16w_kernel<GGB_TYPE_Q4_0><<<grid, kGemmThreads, kGemmSmem, ggb_cuda_stream()>>>(A);
    else gemm_bf16w_kernel<GGB_TYPE_Q8_0><<<grid, kGemmThreads, kGemmSmem, ggb_cuda_stream()>>>(A);
    GGB_LAUNCH_CHECK();
    if (splits > 1) {
        A.resid = resid;
        const int64_t cnt = T * N;
        splitk_reduce_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ggb_cuda_stream()>>>(A, splits);
        GGB_LAUNCH_CHECK();
    }
    return GGB_OK;
}

int ggb_op_gemm_bf16(const ggb_weight *w, const void *a_bf16, int64_t T, float *c, int64_t ldc, const float *resid,
                     void *workspace, size_t workspace_bytes) {
    const ggb_weight *ws[1] = {w};
    float *cs[1] = {c};
    const int64_t ls[1] = {ldc};
    GGB_CHECK_ARG(w != nullptr, "bad gemm args");
    return gemm_launch(1, ws, a_bf16, T, cs, ls, resid, workspace, workspace_bytes);
}

int ggb_op_gemm_bf16_multi(int n, const ggb_weight *const *w, const void *a_bf16, int64_t T, float *const *c, const int64_t *ldc,
                           void *workspace, size_t workspace_bytes) {
    return gemm_launch(n, w, a_bf16, T, c, ldc, nullptr, workspace, workspace_bytes);
}

int ggb_op_attn_prefill(const float *q, const void *kcache, const void *vcache, int64_t T, int pos0, int H, int Hkv, int d,
                        float *out, void *out_bf16) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(q && kcache && vcache && (out || out_bf16) && T > 0 && d % 8 == 0 && d <= 256 && H % Hkv == 0, "bad attn_prefill args");
    const int n_kv_max = pos0 + (int)T;
    if ((H / Hkv) % 4 == 0 && !getenv("GGB_PREFILL_ATTN_SIMPLE")) { // MQA / GQA: one CTA per (token, 4 heads of one KV head)
        const size_t smem4 = ((size_t)4 * n_kv_max + 4 * 4 * 256 + 16) * 4 + 16 * 8;
        GGB_CHECK_ARG(smem4 <= 200 * 1024, "attn_prefill: context %d too long for the score buffer", n_kv_max);
        if (smem4 > 48 * 1024) GGB_CUDA(cudaFuncSetAttribute(attn_prefill_gqa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
        attn_prefill_gqa_kernel<<<(unsigned)(T * (H / 4)), 128, smem4, ggb_cuda_stream()>>>(
            q, (const __half *)kcache, (const __half *)vcache, (int)T, pos0, H, Hkv, d, ggb_exp_table_dev(), out,
            (__nv_bfloat16 *)out_bf16, n_kv_max);
        GGB_LAUNCH_CHECK();
        return GGB_OK;
    }
    // query heads per warp: sharing the K/V loads between 4 heads (gh = 4) measured SLOWER for Gemma-2B at T = 512 (3.9 vs
    // 3.3 ms per prefill: a quarter of the warps, each with 4x the serial work), so one head per warp it is
    const int gh = getenv("GGB_PREFILL_ATTN_GH4") && (H / Hkv) % 4 == 0 ? 4 : 1;
    const size_t smem = (size_t)kAttnWarps * gh * n_kv_max * 4;
    GGB_CHECK_ARG(smem <= 200 * 1024, "attn_prefill: context %d too long for the score buffer", n_kv_max);
    const unsigned blocks = (unsigned)((T * (H / gh) + kAttnWarps - 1) / kAttnWarps);
    if (gh == 4) {
        if (smem > 48 * 1024) GGB_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_prefill_kernel<4><<<blocks, kAttnWarps * 32, smem, ggb_cuda_stream()>>>(q, (const __half *)kcache, (const __half *)vcache,
                                                                                    (int)T, pos0, H, Hkv, d, ggb_exp_table_dev(), out, (__nv_bfloat16 *)out_bf16, n_kv_max);
    } else {
        if (smem > 48 * 1024) GGB_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_prefill_kernel<1><<<blocks, kAttnWarps * 32, smem, ggb_cuda_stream()>>>(q, (const __half *)kcache, (const __half *)vcache,
                                                                                    (int)T, pos0, H, Hkv, d, ggb_exp_table_dev(), out, (__nv_bfloat16 *)out_bf16, n_kv_max);
    }
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

} // extern "C"
