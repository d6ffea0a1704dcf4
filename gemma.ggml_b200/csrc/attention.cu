// attention.cu — K4: MQA/MHA decode attention over the fp16 KV cache, per-operator form.
// Replaces the reference's three graph nodes kq = mul_mat(k, q) -> soft_max_ext -> kqv = mul_mat(v, kq)
// (ref: src/gemma_model.cpp:454-489) with two kernels that reproduce the reference's arithmetic:
//
//   scores  one CTA per (kv head, 32-position chunk): kq[h][p] = sum_i K16[p][i] * fp16(q[h][i]), fp16 x fp16
//           products (exact) accumulated in DOUBLE and rounded to fp32 once, as the reference's vec_dot_f16
//           does (ref: src/kernals.cl:158-164).  All G = H/Hkv query heads of a KV head share the K chunk.
//   pv      one CTA per (head, 16-dim slice): the whole soft_max_ext row of the head ([UPSTREAM]
//           ggml_compute_forward_soft_max: global max, p = exp_table[f16(s - max)], sum in double - exact for
//           fp16-valued terms in any order -, p *= (float)(1/sum)), P rounded to fp16 (ggml mul_mat INIT of the
//           kqv node), then out[i] = sum_p V16[p][i] * P16[p] accumulated in double.
//
// Double accumulators make the sums insensitive to their order (differences ~1e-16 relative, invisible after the
// final rounding to fp32), so the split over lanes / positions needs no combine pass and no second order of its
// own: results are bit-identical to the scalar reference order (tests compare bits).
//
// KV cache layout in HBM: position-major fp16 [n_ctx][Hkv*d] for BOTH K and V (the reference keeps V
// transposed [d][n_ctx], :481-482, which would turn the per-token V store into d scattered 2-byte writes).
#include "common.cuh"

namespace {

constexpr int kChunk = GGB_DEV_ATTN_CHUNK; // 32 positions
constexpr int kMaxG = 8;                   // query heads per KV head handled per pass
constexpr int kSlice = 16;                 // dims per pv CTA

__device__ __forceinline__ int clamp_npos(const int32_t *pos_p, int n_ctx) {
    return min(max(*pos_p, 0), n_ctx - 1) + 1;
}

// grid (max_chunks, Hkv, ceil(G/8)); 256 threads
__global__ void __launch_bounds__(256)
attn_scores_kernel(const float *__restrict__ q, const __half *__restrict__ kcache, const int32_t *__restrict__ pos_p,
                   int H, int Hkv, int d, int n_ctx, float *__restrict__ scores) {
    const int n_pos = clamp_npos(pos_p, n_ctx);
    const int p0 = blockIdx.x * kChunk;
    if (p0 >= n_pos) return;
    const int np = min(kChunk, n_pos - p0);
    const int g = blockIdx.y;
    const int G = H / Hkv;
    const int h0 = g * G + blockIdx.z * kMaxG;
    const int ng = min(kMaxG, G - blockIdx.z * kMaxG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t kv_stride = (int64_t)Hkv * d;
    __shared__ double s_q[kMaxG][256]; // fp16-rounded q, widened (d <= 256)
    for (int i = tid; i < ng * d; i += 256) {
        const int jh = i / d, e = i % d;
        s_q[jh][e] = (double)__half2float(__float2half_rn(q[(int64_t)(h0 + jh) * d + e]));
    }
    __syncthreads();
    const bool in = 8 * lane < d;
#pragma unroll 1
    for (int j = warp; j < np; j += 8) {
        double kd[8];
        if (in) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(kcache + (int64_t)(p0 + j) * kv_stride + (int64_t)g * d + 8 * lane);
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __half22float2(h2[e]);
                kd[2 * e] = (double)f.x;
                kd[2 * e + 1] = (double)f.y;
            }
        }
        for (int jh = 0; jh < ng; jh++) {
            double acc = 0.0;
            if (in) {
#pragma unroll
                for (int e = 0; e < 8; e++) acc = __fma_rn(kd[e], s_q[jh][8 * lane + e], acc);
            }
            acc = warp_sum_f64(acc);
            if (lane == 0) scores[(size_t)(h0 + jh) * n_ctx + p0 + j] = __double2float_rn(acc);
        }
    }
}

// grid (H, d / 16); 256 threads; dynamic shared memory: n_ctx floats
__global__ void __launch_bounds__(256)
attn_pv_kernel(const float *__restrict__ scores, const __half *__restrict__ vcache, const int32_t *__restrict__ pos_p,
               int H, int Hkv, int d, int n_ctx, const uint16_t *__restrict__ exp_table, float *__restrict__ out) {
    extern __shared__ float s_p[]; // scores -> e -> fp16-rounded p
    __shared__ float s_redf[8];
    __shared__ double s_redd[8];
    __shared__ double s_acc[8][16];
    const int n_pos = clamp_npos(pos_p, n_ctx);
    const int h = blockIdx.x, slice = blockIdx.y, g = h / (H / Hkv);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *srow = scores + (size_t)h * n_ctx;
    float mx = -INFINITY;
    for (int i = tid; i < n_pos; i += 256) {
        const float s = srow[i];
        s_p[i] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max_f32(mx);
    if (lane == 0) s_redf[warp] = mx;
    __syncthreads();
    mx = s_redf[0];
#pragma unroll
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, s_redf[w]);
    double sum = 0.0;
    for (int i = tid; i < n_pos; i += 256) {
        const float s = s_p[i];
        float e = 0.0f;
        if (s != -INFINITY) {
            e = half_bits_to_float(exp_table[float_to_half_bits(__fsub_rn(s, mx))]);
            sum = __dadd_rn(sum, (double)e);
        }
        s_p[i] = e;
    }
    sum = warp_sum_f64(sum);
    if (lane == 0) s_redd[warp] = sum;
    __syncthreads();
    sum = s_redd[0];
#pragma unroll
    for (int w = 1; w < 8; w++) sum = __dadd_rn(sum, s_redd[w]);
    const float fs = __double2float_rn(__ddiv_rn(1.0, sum));
    for (int i = tid; i < n_pos; i += 256) s_p[i] = __half2float(__float2half_rn(__fmul_rn(s_p[i], fs)));
    __syncthreads();
    // out[i] = sum_p V16[p][i] * P16[p]: thread (o = dim octet, pl = position lane)
    const int o = tid & 1, pl = tid >> 1;
    const __half *vbase = vcache + (size_t)g * d + (size_t)slice * kSlice + 8 * o;
    const size_t kv_stride = (size_t)Hkv * d;
    double acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.0;
    for (int p = pl; p < n_pos; p += 128) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(vbase + (size_t)p * kv_stride);
        const double pd = (double)s_p[p];
        const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float2 f = __half22float2(h2[e]);
            acc[2 * e] = __fma_rn((double)f.x, pd, acc[2 * e]);
            acc[2 * e + 1] = __fma_rn((double)f.y, pd, acc[2 * e + 1]);
        }
    }
#pragma unroll
    for (int off = 2; off <= 16; off <<= 1) {
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = __dadd_rn(acc[e], __shfl_xor_sync(0xffffffffu, acc[e], off));
    }
    if (lane < 2) {
#pragma unroll
        for (int e = 0; e < 8; e++) s_acc[warp][8 * lane + e] = acc[e];
    }
    __syncthreads();
    if (tid < kSlice) {
        double t = s_acc[0][tid];
#pragma unroll
        for (int w = 1; w < 8; w++) t = __dadd_rn(t, s_acc[w][tid]);
        out[(size_t)h * d + (size_t)slice * kSlice + tid] = __double2float_rn(t);
    }
}

} // namespace

extern "C" {

// scores [H][n_ctx] + a [H*d] fp32 row (kqv when the caller only wants its Q8_0 form)
size_t ggb_attn_scratch_bytes(int H, int d, int n_ctx) { return sizeof(float) * ((size_t)H * n_ctx + (size_t)H * d); }

int ggb_op_attn_decode(const float *q, const void *kcache, const void *vcache, const int32_t *pos, int H, int Hkv,
                       int d, int n_ctx, void *scratch, float *out, void *q8_out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(q && kcache && vcache && pos && scratch && (out || q8_out), "null pointer");
    GGB_CHECK_ARG(Hkv > 0 && H % Hkv == 0, "n_head %d not a multiple of n_head_kv %d", H, Hkv);
    GGB_CHECK_ARG(d % 32 == 0 && d <= 256, "head_dim %d unsupported (multiple of 32, <= 256)", d);
    GGB_CHECK_ARG(n_ctx > 0 && (size_t)n_ctx * 4 <= 200 * 1024, "n_ctx %d unsupported", n_ctx);
    float *scores = (float *)scratch;
    float *kqv = out ? out : scores + (size_t)H * n_ctx;
    const int G = H / Hkv;
    const int max_chunks = (n_ctx + kChunk - 1) / kChunk;
    const dim3 grid((unsigned)max_chunks, (unsigned)Hkv, (unsigned)((G + kMaxG - 1) / kMaxG));
    attn_scores_kernel<<<grid, 256, 0, ggb_cuda_stream()>>>(q, (const __half *)kcache, pos, H, Hkv, d, n_ctx, scores);
    GGB_LAUNCH_CHECK();
    const size_t smem = sizeof(float) * (size_t)n_ctx;
    if (smem > 48 * 1024)
        GGB_CUDA(cudaFuncSetAttribute(attn_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_pv_kernel<<<dim3((unsigned)H, (unsigned)(d / kSlice)), 256, smem, ggb_cuda_stream()>>>(
        scores, (const __half *)vcache, pos, H, Hkv, d, n_ctx, ggb_exp_table_dev(), kqv);
    GGB_LAUNCH_CHECK();
    if (q8_out) return ggb_op_quantize_q8_0(kqv, (int64_t)H * d, 1, q8_out);
    return GGB_OK;
}

} // extern "C"
