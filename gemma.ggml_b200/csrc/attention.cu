// attention.cu — K4: MQA/MHA decode attention over the fp16 KV cache, split-K across the context.
// Replaces the reference's three graph nodes kq = mul_mat(k, q) -> soft_max_ext -> kqv = mul_mat(v, kq)
// (ref: src/gemma_model.cpp:454-489) with one partial kernel per (kv-head, 32-position chunk) and one
// combine kernel.  All G = H/Hkv query heads that share a KV head are processed against the same
// K/V chunk, so K and V are read from HBM exactly once per token (MQA: G = 8).
//
// Numerics (DEV order, mirrored by oracle orc_attn_decode_dev; tolerance vs the reference's
// two-pass fp16-P form is stated in tests/): q is rounded to fp16 as ggml's mul_mat INIT does;
// scores accumulate in fp32 (lane l: elements 8l..8l+7, then butterfly); per-chunk softmax through
// the fp16 exp table; P·V sequential over the chunk's positions; chunks combined in order.
//
// KV cache layout in HBM: position-major fp16 [n_ctx][Hkv*d] for BOTH K and V (the reference keeps V
// transposed [d][n_ctx], :481-482, which would turn the per-token V store into d scattered 2-byte
// writes and the V read into strided gathers).
#include "common.cuh"

namespace {

constexpr int kChunk = GGB_DEV_ATTN_CHUNK; // 32 positions
constexpr int kMaxG = 8;                   // query heads per KV head handled per pass

// scratch layout per head h, chunk c: m[h][c], l[h][c], acc[h][c][d]
struct AttnScratch {
    float *m, *l, *acc;
    int max_chunks;
};

__host__ __device__ inline AttnScratch carve_scratch(void *scratch, int H, int d, int n_ctx) {
    AttnScratch s;
    s.max_chunks = (n_ctx + kChunk - 1) / kChunk;
    s.m = (float *)scratch;
    s.l = s.m + (size_t)H * s.max_chunks;
    s.acc = s.l + (size_t)H * s.max_chunks;
    return s;
}

// grid (max_chunks, Hkv, ceil(G/8)); 256 threads
__global__ void __launch_bounds__(256)
attn_partial_kernel(const float *__restrict__ q, const __half *__restrict__ kcache, const __half *__restrict__ vcache,
                    const int32_t *__restrict__ pos_p, int H, int Hkv, int d, AttnScratch sc,
                    const uint16_t *__restrict__ exp_table) {
    const int n_pos = *pos_p + 1;
    const int chunk = blockIdx.x;
    const int p0 = chunk * kChunk;
    if (p0 >= n_pos) return;
    const int np = min(kChunk, n_pos - p0);
    const int g = blockIdx.y;
    const int G = H / Hkv;
    const int h0 = g * G + blockIdx.z * kMaxG; // first query head of this pass
    const int ng = min(kMaxG, G - blockIdx.z * kMaxG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t kv_stride = (int64_t)Hkv * d;

    __shared__ float s_q[kMaxG][256]; // fp16-rounded q, widened (d <= 256)
    __shared__ float s_s[kMaxG][kChunk];
    __shared__ float s_e[kMaxG][kChunk];

    for (int i = tid; i < ng * d; i += 256) {
        const int jh = i / d, e = i % d;
        s_q[jh][e] = __half2float(__float2half_rn(q[(int64_t)(h0 + jh) * d + e]));
    }
    __syncthreads();

    // phase 1: scores.  warp w takes positions w, w+8, w+16, w+24 of the chunk
    {
        uint4 kraw[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int j = warp + 8 * i;
            if (j < np && 8 * lane < d)
                kraw[i] = *reinterpret_cast<const uint4 *>(kcache + (int64_t)(p0 + j) * kv_stride + (int64_t)g * d + 8 * lane);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int j = warp + 8 * i;
            if (j < np) { // warp-uniform
                float kf[8];
                if (8 * lane < d) {
                    const __half2 *h2 = reinterpret_cast<const __half2 *>(&kraw[i]);
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 f = __half22float2(h2[e]);
                        kf[2 * e] = f.x;
                        kf[2 * e + 1] = f.y;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) kf[e] = 0.f;
                }
                for (int jh = 0; jh < ng; jh++) {
                    float acc = 0.0f;
                    if (8 * lane < d) {
#pragma unroll
                        for (int e = 0; e < 8; e++) acc = __fadd_rn(acc, __fmul_rn(kf[e], s_q[jh][8 * lane + e]));
                    }
                    acc = warp_sum_f32(acc);
                    if (lane == 0) s_s[jh][j] = acc;
                }
            }
        }
    }
    __syncthreads();

    // phase 2: per-head chunk softmax.  warp w handles head w; lane = position
    if (warp < ng) {
        const float sv = lane < np ? s_s[warp][lane] : -INFINITY;
        const float m = warp_max_f32(sv);
        const float e = lane < np ? half_bits_to_float(exp_table[float_to_half_bits(__fsub_rn(sv, m))]) : 0.0f;
        const float l = warp_sum_f32(e);
        s_e[warp][lane] = e;
        if (lane == 0) {
            sc.m[(size_t)(h0 + warp) * sc.max_chunks + chunk] = m;
            sc.l[(size_t)(h0 + warp) * sc.max_chunks + chunk] = l;
        }
    }
    __syncthreads();

    // phase 3: acc[h][i] = sum_p e[h][p] * V[p][i], sequential in p.  thread -> (dim pair, head parity)
    {
        const int halfd = d / 2;
        const int dp = tid % 128, hs = tid / 128; // 128 dim pairs max (d <= 256), 2 head subsets
        if (dp < halfd) {
            float a0[kMaxG / 2], a1[kMaxG / 2];
#pragma unroll
            for (int u = 0; u < kMaxG / 2; u++) a0[u] = a1[u] = 0.0f;
            for (int j = 0; j < np; j++) {
                const __half2 v2 = *reinterpret_cast<const __half2 *>(vcache + (int64_t)(p0 + j) * kv_stride +
                                                                     (int64_t)g * d + 2 * dp);
                const float2 vf = __half22float2(v2);
#pragma unroll
                for (int u = 0; u < kMaxG / 2; u++) {
                    const int jh = 2 * u + hs;
                    if (jh < ng) {
                        const float e = s_e[jh][j];
                        a0[u] = __fadd_rn(a0[u], __fmul_rn(e, vf.x));
                        a1[u] = __fadd_rn(a1[u], __fmul_rn(e, vf.y));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kMaxG / 2; u++) {
                const int jh = 2 * u + hs;
                if (jh < ng) {
                    float *out = sc.acc + ((size_t)(h0 + jh) * sc.max_chunks + chunk) * d + 2 * dp;
                    *reinterpret_cast<float2 *>(out) = make_float2(a0[u], a1[u]);
                }
            }
        }
    }
}

// combine: grid H, d threads.  M = max m_c; w_c = exp_table[f16(m_c - M)]; L = sum l_c*w_c;
// O[i] = sum acc_c[i]*w_c (sequential in c); out = O / L; optional Q8_0 quantisation of the
// row for the output projection (ggml mul_mat INIT of kqv_out).
__global__ void __launch_bounds__(256)
attn_combine_kernel(const int32_t *__restrict__ pos_p, int H, int d, AttnScratch sc,
                    const uint16_t *__restrict__ exp_table, float *__restrict__ out, uint8_t *__restrict__ q8_out) {
    extern __shared__ float s_w[]; // [nchunk] weights, then [nchunk] l
    __shared__ float s_red[8];
    const int n_pos = *pos_p + 1;
    const int nchunk = (n_pos + kChunk - 1) / kChunk;
    const int h = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *pm = sc.m + (size_t)h * sc.max_chunks;
    const float *pl = sc.l + (size_t)h * sc.max_chunks;
    float mx = -INFINITY;
    for (int c = tid; c < nchunk; c += blockDim.x) mx = fmaxf(mx, pm[c]);
    mx = warp_max_f32(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) mx = fmaxf(mx, s_red[w]);
    for (int c = tid; c < nchunk; c += blockDim.x) {
        s_w[c] = half_bits_to_float(exp_table[float_to_half_bits(__fsub_rn(pm[c], mx))]);
        s_w[nchunk + c] = pl[c];
    }
    __syncthreads();
    float L = 0.0f, O = 0.0f;
    const float *pacc = sc.acc + (size_t)h * sc.max_chunks * d + tid;
    for (int c = 0; c < nchunk; c++) {
        const float w = s_w[c];
        L = __fadd_rn(L, __fmul_rn(s_w[nchunk + c], w));
        if (tid < d) O = __fadd_rn(O, __fmul_rn(pacc[(size_t)c * d], w));
    }
    const float r = __fdiv_rn(O, L);
    if (tid < d && out) out[(size_t)h * d + tid] = r;
    if (q8_out) { // d % 32 == 0: each warp's 32 outputs are one Q8_0 block of the [H*d] row
        float dd;
        const int qv = quantize_q8_lane(tid < d ? r : 0.0f, dd);
        if (tid < d) {
            ((int8_t *)q8_out)[(size_t)h * d + tid] = (int8_t)qv;
            if (lane == 0)
                reinterpret_cast<__half *>(q8_out + (size_t)H * d)[((size_t)h * d + tid) >> 5] = __float2half_rn(dd);
        }
    }
}

} // namespace

extern "C" {

size_t ggb_attn_scratch_bytes(int H, int d, int n_ctx) {
    const size_t chunks = (size_t)(n_ctx + kChunk - 1) / kChunk;
    return sizeof(float) * (size_t)H * chunks * (2 + (size_t)d);
}

int ggb_op_attn_decode(const float *q, const void *kcache, const void *vcache, const int32_t *pos, int H, int Hkv,
                       int d, int n_ctx, void *scratch, float *out, void *q8_out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(q && kcache && vcache && pos && scratch && (out || q8_out), "null pointer");
    GGB_CHECK_ARG(Hkv > 0 && H % Hkv == 0, "n_head %d not a multiple of n_head_kv %d", H, Hkv);
    GGB_CHECK_ARG(d % 32 == 0 && d <= 256, "head_dim %d unsupported (multiple of 32, <= 256)", d);
    const AttnScratch sc = carve_scratch(scratch, H, d, n_ctx);
    const int G = H / Hkv;
    const dim3 grid((unsigned)sc.max_chunks, (unsigned)Hkv, (unsigned)((G + kMaxG - 1) / kMaxG));
    attn_partial_kernel<<<grid, 256, 0, ggb_cuda_stream()>>>(q, (const __half *)kcache, (const __half *)vcache, pos, H,
                                                            Hkv, d, sc, ggb_exp_table_dev());
    GGB_LAUNCH_CHECK();
    const int threads = d < 32 ? 32 : d;
    const size_t smem = sizeof(float) * 2 * (size_t)sc.max_chunks;
    attn_combine_kernel<<<(unsigned)H, threads, smem, ggb_cuda_stream()>>>(pos, H, d, sc, ggb_exp_table_dev(), out,
                                                                          (uint8_t *)q8_out);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

} // extern "C"
