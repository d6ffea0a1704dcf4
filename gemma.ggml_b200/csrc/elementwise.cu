// elementwise.cu — K1 quantize_row_q8_0, K5 rms_norm(+mul)(+quantize), K6 rope NEOX (+Q scale,
// +KV-cache store), K8 get_rows+scale, K9 argmax, and the small element-wise graph nodes
// (add / mul / scale / gelu / soft_max_ext / cpy) the builder emits
// (ref: src/gemma_model.cpp:438-452, 499-518, 675-679, 698-716; [UPSTREAM] ggml op semantics,
// SURVEY.md Appendix A).  All HBM-bound or latency-bound; no tensor cores.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- K1 quantize_row_q8_0
// one warp per 32-float block, lane = element ([UPSTREAM] quantize_row_q8_0_reference:
// d = amax/127, id = 1/d, q = roundf(x*id)); bit-exact by construction (IEEE div, roundf).
__global__ void quantize_q8_kernel(const float *__restrict__ x, int64_t cols, int64_t T, uint8_t *__restrict__ q8,
                                   int64_t q8_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t blk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nb = cols / GGB_QK;
    if (blk >= nb * T) return;
    const int64_t t = blk / nb, b = blk % nb;
    float d;
    const int q = quantize_q8_lane(x[t * cols + b * 32 + lane], d);
    uint8_t *row = q8 + t * q8_stride;
    ((int8_t *)row)[b * 32 + lane] = (int8_t)q;
    if (lane == 0) reinterpret_cast<__half *>(row + cols)[b] = __float2half_rn(d);
}

__global__ void q8_to_blocks_kernel(const uint8_t *__restrict__ q8, int64_t cols, int64_t T, int64_t q8_stride,
                                    uint8_t *__restrict__ blocks) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nb = cols / GGB_QK;
    if (idx >= nb * T) return;
    const int64_t t = idx / nb, b = idx % nb;
    const uint8_t *row = q8 + t * q8_stride;
    uint8_t *out = blocks + idx * 34;
    const uint16_t dbits = __half_as_ushort(reinterpret_cast<const __half *>(row + cols)[b]);
    out[0] = (uint8_t)(dbits & 0xff);
    out[1] = (uint8_t)(dbits >> 8);
    for (int i = 0; i < 32; i++) out[2 + i] = row[b * 32 + i];
}

__global__ void q8_from_blocks_kernel(const uint8_t *__restrict__ blocks, int64_t cols, int64_t T, int64_t row_size,
                                      int64_t q8_stride, uint8_t *__restrict__ q8) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nb = cols / GGB_QK;
    if (idx >= nb * T) return;
    const int64_t t = idx / nb, b = idx % nb;
    const uint8_t *in = blocks + t * row_size + b * 34;
    uint8_t *row = q8 + t * q8_stride;
    reinterpret_cast<__half *>(row + cols)[b] = __ushort_as_half((uint16_t)in[0] | ((uint16_t)in[1] << 8));
    for (int i = 0; i < 32; i++) row[b * 32 + i] = in[2 + i];
}

// ---------------------------------------------------------------- K5 rms_norm (+mul, +quantize)
// [UPSTREAM] ggml_compute_forward_rms_norm_f32 then ggml_mul (ref: src/gemma_model.cpp:438-442).
// Sum of squares in double (ggml_float) of the fp32-rounded products: 256 strided partials, butterfly, warps in order - the
// grouping of a double sum is invisible after the rounding to fp32, so this is the scalar loop's result.
__global__ void __launch_bounds__(GGB_DEV_NORM_THREADS)
rms_norm_kernel(const float *__restrict__ x, const float *__restrict__ w, float eps, int64_t E, float *__restrict__ y,
                uint8_t *__restrict__ q8, int64_t q8_stride) {
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
    for (int64_t i = tid; i < E; i += GGB_DEV_NORM_THREADS) {
        float v = __fmul_rn(xr[i], scale);
        if (w) v = __fmul_rn(v, w[i]);
        if (y) y[(int64_t)blockIdx.x * E + i] = v;
        if (q8) { // E % 32 == 0: the warp's 32 consecutive elements are one Q8_0 block
            float d;
            const int q = quantize_q8_lane(v, d);
            uint8_t *row = q8 + (int64_t)blockIdx.x * q8_stride;
            ((int8_t *)row)[i] = (int8_t)q;
            if (lane == 0) reinterpret_cast<__half *>(row + E)[i >> 5] = __float2half_rn(d);
        }
    }
}

// ---------------------------------------------------------------- element-wise nodes
__global__ void add_kernel(const float *a, const float *b, int64_t n, float *y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __fadd_rn(a[i], b[i]);
}
__global__ void mul_kernel(const float *a, const float *b, int64_t n, float *y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __fmul_rn(a[i], b[i]);
}
__global__ void scale_kernel(const float *x, float s, int64_t n, float *y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __fmul_rn(x[i], s);
}
// [UPSTREAM] ggml_vec_gelu_f32 (GGML_GELU_FP16): y = table[f16(x)]
__global__ void gelu_kernel(const float *x, int64_t n, const uint16_t *__restrict__ table, float *y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = half_bits_to_float(table[float_to_half_bits(x[i])]);
}
__global__ void cpy_f32_f16_kernel(const float *x, int64_t n, __half *y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2half_rn(x[i]);
}

// [UPSTREAM] ggml_compute_forward_soft_max_f32 (max_bias 0): w = x*scale + mask; p = exp_table[f16(w - max)]
// (0 where w == -inf); sum in double; p *= (float)(1/sum).  One CTA per row.
__global__ void __launch_bounds__(256)
soft_max_kernel(const float *__restrict__ x, const float *__restrict__ mask, float scale, int64_t nc,
                const uint16_t *__restrict__ table, float *__restrict__ y) {
    __shared__ float s_m[8];
    __shared__ double s_s[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *xr = x + (int64_t)blockIdx.x * nc;
    const float *mr = mask ? mask + (int64_t)blockIdx.x * nc : nullptr;
    float *yr = y + (int64_t)blockIdx.x * nc;
    float mx = -INFINITY;
    for (int64_t i = tid; i < nc; i += 256) {
        float wv = __fmul_rn(xr[i], scale);
        if (mr) wv = __fadd_rn(wv, mr[i]);
        mx = fmaxf(mx, wv);
    }
    mx = warp_max_f32(mx);
    if (lane == 0) s_m[warp] = mx;
    __syncthreads();
    mx = s_m[0];
#pragma unroll
    for (int i = 1; i < 8; i++) mx = fmaxf(mx, s_m[i]);
    double acc = 0.0;
    for (int64_t i = tid; i < nc; i += 256) {
        float wv = __fmul_rn(xr[i], scale);
        if (mr) wv = __fadd_rn(wv, mr[i]);
        float p = 0.0f;
        if (wv != -INFINITY) {
            p = half_bits_to_float(table[float_to_half_bits(__fsub_rn(wv, mx))]);
            acc = __dadd_rn(acc, (double)p);
        }
        yr[i] = p;
    }
    acc = warp_sum_f64(acc);
    if (lane == 0) s_s[warp] = acc;
    __syncthreads();
    double sum = s_s[0];
#pragma unroll
    for (int i = 1; i < 8; i++) sum = __dadd_rn(sum, s_s[i]);
    const float inv = __double2float_rn(__ddiv_rn(1.0, sum));
    for (int64_t i = tid; i < nc; i += 256) yr[i] = __fmul_rn(yr[i], inv);
}

// ---------------------------------------------------------------- K6 rope NEOX
// [UPSTREAM] rope mode 2: pairs (i, i + n_dims/2); cos/sin from the host-built recurrence table.
__global__ void rope_kernel(const float *__restrict__ x, const int32_t *__restrict__ pos, int n_dims, int n_heads,
                            int64_t T, const float *__restrict__ cosT, const float *__restrict__ sinT, int table_ctx,
                            float *__restrict__ y) {
    const int half = n_dims / 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * n_heads * half) return;
    const int i = (int)(idx % half);
    const int64_t th = idx / half; // token*n_heads + head
    const int64_t t = th / n_heads;
    int p = pos[t];
    p = p < 0 ? 0 : (p >= table_ctx ? table_ctx - 1 : p);
    const float c = cosT[(int64_t)p * half + i], s = sinT[(int64_t)p * half + i];
    const float x0 = x[th * n_dims + i], x1 = x[th * n_dims + i + half];
    y[th * n_dims + i] = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
    y[th * n_dims + i + half] = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
}

// fused decode form: rope(q)*q_scale in place; rope(k) -> f16 cache row; v -> f16 cache row
// (ref: src/gemma_model.cpp:698-716 rope+scale, :499-518 KV store as two ggml_cpy F32->F16).
__global__ void rope_kv_store_kernel(float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                                     const int32_t *__restrict__ pos_p, int H, int Hkv, int d, float q_scale,
                                     const float *__restrict__ cosT, const float *__restrict__ sinT, int table_ctx,
                                     __half *__restrict__ kcache, __half *__restrict__ vcache) {
    const int half = d / 2;
    int pos = *pos_p; // clamped like the batch form: a position outside the rope table / cache must not address memory
    pos = pos < 0 ? 0 : (pos >= table_ctx ? table_ctx - 1 : pos);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nq = H * half, nk = Hkv * half, nv = Hkv * d;
    if (idx < nq + nk) {
        const bool is_q = idx < nq;
        const int j = is_q ? idx : idx - nq;
        const int h = j / half, i = j % half;
        const float c = cosT[(int64_t)pos * half + i], s = sinT[(int64_t)pos * half + i];
        const float *src = is_q ? q : k;
        const float x0 = src[h * d + i], x1 = src[h * d + i + half];
        const float r0 = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        const float r1 = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
        if (is_q) {
            q[h * d + i] = __fmul_rn(r0, q_scale);
            q[h * d + i + half] = __fmul_rn(r1, q_scale);
        } else {
            __half *row = kcache + (int64_t)pos * Hkv * d;
            row[h * d + i] = __float2half_rn(r0);
            row[h * d + i + half] = __float2half_rn(r1);
        }
    } else if (idx < nq + nk + nv) {
        const int j = idx - nq - nk;
        vcache[(int64_t)pos * Hkv * d + j] = __float2half_rn(v[j]);
    }
}

// the same for T tokens (prefill): token t = blockIdx.y, its position pos[t]; q / k / v are [T][H*d] / [T][Hkv*d]
__global__ void rope_kv_store_batch_kernel(float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                                           const int32_t *__restrict__ pos_p, int H, int Hkv, int d, float q_scale,
                                           const float *__restrict__ cosT, const float *__restrict__ sinT, int table_ctx,
                                           __half *__restrict__ kcache, __half *__restrict__ vcache) {
    const int half = d / 2;
    const int64_t t = blockIdx.y;
    int pos = pos_p[t];
    pos = pos < 0 ? 0 : (pos >= table_ctx ? table_ctx - 1 : pos);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nq = H * half, nk = Hkv * half, nv = Hkv * d;
    float *qt = q + t * (int64_t)H * d;
    const float *kt = k + t * (int64_t)Hkv * d, *vt = v + t * (int64_t)Hkv * d;
    if (idx < nq + nk) {
        const bool is_q = idx < nq;
        const int j = is_q ? idx : idx - nq;
        const int h = j / half, i = j % half;
        const float c = cosT[(int64_t)pos * half + i], s = sinT[(int64_t)pos * half + i];
        const float *src = is_q ? qt : kt;
        const float x0 = src[h * d + i], x1 = src[h * d + i + half];
        const float r0 = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, s));
        const float r1 = __fadd_rn(__fmul_rn(x0, s), __fmul_rn(x1, c));
        if (is_q) {
            qt[h * d + i] = __fmul_rn(r0, q_scale);
            qt[h * d + i + half] = __fmul_rn(r1, q_scale);
        } else {
            __half *row = kcache + (int64_t)pos * Hkv * d;
            row[h * d + i] = __float2half_rn(r0);
            row[h * d + i + half] = __float2half_rn(r1);
        }
    } else if (idx < nq + nk + nv) {
        const int j = idx - nq - nk;
        vcache[(int64_t)pos * Hkv * d + j] = __float2half_rn(vt[j]);
    }
}

// ---------------------------------------------------------------- K8 get_rows + scale
// [UPSTREAM] dequantize_row_q4_0 / q8_0 of the selected row, then ggml_scale (ref: :675-679).
__global__ void get_rows_kernel(const uint8_t *__restrict__ qs, const __half *__restrict__ d, int type, int64_t nb,
                                int64_t n_rows, const int32_t *__restrict__ ids, int64_t id_stride, float scale,
                                float *__restrict__ y) {
    const int64_t t = blockIdx.y;
    int64_t row = ids[t * id_stride];
    row = row < 0 ? 0 : (row >= n_rows ? n_rows - 1 : row);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // element in row
    if (e >= nb * 32) return;
    const int64_t b = e >> 5;
    const int j = (int)(e & 31);
    const float dd = __half2float(d[row * nb + b]);
    float val;
    if (type == GGB_TYPE_Q4_0) {
        const uint8_t byte = qs[(row * nb + b) * 16 + (j & 15)];
        const int qv = (j < 16 ? (byte & 0x0F) : (byte >> 4)) - 8;
        val = __fmul_rn((float)qv, dd);
    } else {
        val = __fmul_rn((float)((const int8_t *)qs)[(row * nb + b) * 32 + j], dd);
    }
    y[t * nb * 32 + e] = __fmul_rn(val, scale);
}

// ---------------------------------------------------------------- K9 argmax (first maximum)
struct Cand {
    float v;
    int i;
};
__device__ __forceinline__ Cand better(Cand a, Cand b) {
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ Cand block_best(Cand c, Cand *s_c) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        Cand o;
        o.v = __shfl_xor_sync(0xffffffffu, c.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, c.i, off);
        c = better(c, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_c[warp] = c;
    __syncthreads();
    Cand r = s_c[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = better(r, s_c[w]);
    return r;
}
// stage 1: each CTA scans a contiguous slice with the reference's strict '>' (ref: :538-543)
__global__ void __launch_bounds__(256) argmax_stage1(const float *__restrict__ x, int64_t n, Cand *__restrict__ part) {
    __shared__ Cand s_c[8];
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per;
    const int64_t hi = lo + per < n ? lo + per : n;
    Cand c{-INFINITY, -1};
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float v = x[i];
        if (v > c.v) {
            c.v = v;
            c.i = (int)i;
        }
    }
    c = block_best(c, s_c);
    if (threadIdx.x == 0) part[blockIdx.x] = c;
}
// stage 2: pick the winner; optionally advance the decode state
__global__ void __launch_bounds__(256)
argmax_stage2(const Cand *__restrict__ part, int n_part, int32_t *__restrict__ out, int32_t *__restrict__ state,
              int32_t *__restrict__ history) {
    __shared__ Cand s_c[8];
    Cand c{-INFINITY, -1};
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) c = better(c, part[i]);
    c = block_best(c, s_c);
    if (threadIdx.x == 0) {
        if (out) *out = c.i;
        if (state) {
            // state = [token, pos, pos_step, n_ctx]: the position stops at n_ctx - 1 (history holds n_ctx + 2 ids)
            int np = state[1] + state[2];
            if (state[3] > 0 && np > state[3] - 1) np = state[3] - 1;
            if (np < 0) np = 0;
            state[0] = c.i;
            state[1] = np;
            if (history) history[np] = c.i;
        }
    }
}

// stage 2 for a row shard: publish (value, global index) so ranks can be compared after an allgather
__global__ void __launch_bounds__(256)
argmax_stage2_shard(const Cand *__restrict__ part, int n_part, int index_offset, Cand *__restrict__ out) {
    __shared__ Cand s_c[8];
    Cand c{-INFINITY, -1};
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) c = better(c, part[i]);
    c = block_best(c, s_c);
    if (threadIdx.x == 0) {
        if (c.i >= 0) c.i += index_offset;
        *out = c;
    }
}

constexpr int kArgmaxParts = 296;

inline unsigned blocks_for(int64_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

} // namespace

extern "C" {

int ggb_op_quantize_q8_0(const float *x, int64_t cols, int64_t T, void *q8) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && q8 && cols > 0 && cols % GGB_QK == 0 && T > 0, "bad quantize args (cols %lld)", (long long)cols);
    const int64_t threads = cols / GGB_QK * T * 32;
    quantize_q8_kernel<<<blocks_for(threads), 256, 0, ggb_cuda_stream()>>>(x, cols, T, (uint8_t *)q8,
                                                                          (int64_t)ggb_q8_row_bytes(cols));
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_q8_to_blocks(const void *q8, int64_t cols, int64_t T, void *blocks) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(q8 && blocks && cols % GGB_QK == 0, "bad args");
    q8_to_blocks_kernel<<<blocks_for(cols / GGB_QK * T), 256, 0, ggb_cuda_stream()>>>(
        (const uint8_t *)q8, cols, T, (int64_t)ggb_q8_row_bytes(cols), (uint8_t *)blocks);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_q8_from_blocks(const void *blocks, int64_t cols, int64_t T, int64_t row_size, void *q8) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(q8 && blocks && cols % GGB_QK == 0 && row_size >= cols / GGB_QK * 34, "bad args");
    q8_from_blocks_kernel<<<blocks_for(cols / GGB_QK * T), 256, 0, ggb_cuda_stream()>>>(
        (const uint8_t *)blocks, cols, T, row_size, (int64_t)ggb_q8_row_bytes(cols), (uint8_t *)q8);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_rms_norm(const float *x, const float *w, float eps, int64_t E, int64_t T, float *y, void *q8) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && E > 0 && T > 0 && (y || q8), "bad rms_norm args");
    GGB_CHECK_ARG(!q8 || E % GGB_QK == 0, "rms_norm+quantize needs E %% 32 == 0");
    rms_norm_kernel<<<(unsigned)T, GGB_DEV_NORM_THREADS, 0, ggb_cuda_stream()>>>(x, w, eps, E, y, (uint8_t *)q8,
                                                                                 (int64_t)ggb_q8_row_bytes(E));
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_add(const float *a, const float *b, int64_t n, float *y) {
    GGB_REQUIRE_INIT();
    add_kernel<<<blocks_for(n), 256, 0, ggb_cuda_stream()>>>(a, b, n, y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}
int ggb_op_mul(const float *a, const float *b, int64_t n, float *y) {
    GGB_REQUIRE_INIT();
    mul_kernel<<<blocks_for(n), 256, 0, ggb_cuda_stream()>>>(a, b, n, y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}
int ggb_op_scale(const float *x, float s, int64_t n, float *y) {
    GGB_REQUIRE_INIT();
    scale_kernel<<<blocks_for(n), 256, 0, ggb_cuda_stream()>>>(x, s, n, y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}
int ggb_op_gelu(const float *x, int64_t n, float *y) {
    GGB_REQUIRE_INIT();
    gelu_kernel<<<blocks_for(n), 256, 0, ggb_cuda_stream()>>>(x, n, ggb_gelu_table_dev(), y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}
int ggb_op_cpy_f32_f16(const float *x, int64_t n, void *y_f16) {
    GGB_REQUIRE_INIT();
    cpy_f32_f16_kernel<<<blocks_for(n), 256, 0, ggb_cuda_stream()>>>(x, n, (__half *)y_f16);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}
int ggb_op_soft_max_ext(const float *x, const float *mask, float scale, int64_t nc, int64_t nrows, float *y) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && y && nc > 0 && nrows > 0, "bad soft_max args");
    soft_max_kernel<<<(unsigned)nrows, 256, 0, ggb_cuda_stream()>>>(x, mask, scale, nc, ggb_exp_table_dev(), y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_rope_neox(const float *x, const int32_t *pos, int n_dims, int n_heads, int64_t T, float *y) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(ggb_rope_cos_dev() && ggb_rope_n_dims() == n_dims, "rope table not initialised for n_dims %d", n_dims);
    rope_kernel<<<blocks_for(T * n_heads * (n_dims / 2)), 256, 0, ggb_cuda_stream()>>>(
        x, pos, n_dims, n_heads, T, ggb_rope_cos_dev(), ggb_rope_sin_dev(), ggb_rope_n_ctx(), y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_rope_kv_store(float *q, const float *k, const float *v, const int32_t *pos, int H, int Hkv, int d,
                         float q_scale, void *kcache, void *vcache) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(ggb_rope_cos_dev() && ggb_rope_n_dims() == d, "rope table not initialised for head_dim %d", d);
    const int total = (H + Hkv) * (d / 2) + Hkv * d;
    rope_kv_store_kernel<<<blocks_for(total), 256, 0, ggb_cuda_stream()>>>(
        q, k, v, pos, H, Hkv, d, q_scale, ggb_rope_cos_dev(), ggb_rope_sin_dev(), ggb_rope_n_ctx(), (__half *)kcache,
        (__half *)vcache);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_rope_kv_store_batch(float *q, const float *k, const float *v, const int32_t *pos, int64_t T, int H, int Hkv, int d,
                               float q_scale, void *kcache, void *vcache) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(q && k && v && pos && kcache && vcache && T > 0 && T <= 65535, "bad rope_kv_store_batch args");
    GGB_CHECK_ARG(ggb_rope_cos_dev() && ggb_rope_n_dims() == d, "rope table not initialised for head_dim %d", d);
    const int total = (H + Hkv) * (d / 2) + Hkv * d;
    const dim3 grid((unsigned)((total + 255) / 256), (unsigned)T);
    rope_kv_store_batch_kernel<<<grid, 256, 0, ggb_cuda_stream()>>>(q, k, v, pos, H, Hkv, d, q_scale, ggb_rope_cos_dev(),
                                                                     ggb_rope_sin_dev(), ggb_rope_n_ctx(), (__half *)kcache,
                                                                     (__half *)vcache);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_get_rows(const ggb_weight *table, const int32_t *ids, int64_t T, float scale, float *y) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(table && ids && y && T > 0, "bad get_rows args");
    const dim3 grid(blocks_for(table->cols), (unsigned)T);
    get_rows_kernel<<<grid, 256, 0, ggb_cuda_stream()>>>(table->qs, table->d, table->type, table->nb, table->rows, ids, 1,
                                                        scale, y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_embed_token(const ggb_weight *table, const int32_t *state, float scale, float *y) {
    return ggb_op_get_rows(table, state, 1, scale, y);
}

size_t ggb_sample_scratch_bytes(int64_t n) {
    (void)n;
    return sizeof(Cand) * kArgmaxParts;
}

int ggb_op_sample_greedy(const float *logits, int64_t n, int32_t *state, int32_t *history, void *scratch) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(logits && state && scratch && n > 0, "bad sample args");
    argmax_stage1<<<kArgmaxParts, 256, 0, ggb_cuda_stream()>>>(logits, n, (Cand *)scratch);
    GGB_LAUNCH_CHECK();
    argmax_stage2<<<1, 256, 0, ggb_cuda_stream()>>>((const Cand *)scratch, kArgmaxParts, nullptr, state, history);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_argmax_shard(const float *x, int64_t n, int32_t index_offset, void *cand_out, void *scratch) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && cand_out && scratch && n > 0, "bad argmax_shard args");
    argmax_stage1<<<kArgmaxParts, 256, 0, ggb_cuda_stream()>>>(x, n, (Cand *)scratch);
    GGB_LAUNCH_CHECK();
    argmax_stage2_shard<<<1, 256, 0, ggb_cuda_stream()>>>((const Cand *)scratch, kArgmaxParts, index_offset, (Cand *)cand_out);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_sample_from_candidates(const void *cands, int n_cands, int32_t *state, int32_t *history) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(cands && state && n_cands > 0, "bad sample_from_candidates args");
    argmax_stage2<<<1, 256, 0, ggb_cuda_stream()>>>((const Cand *)cands, n_cands, nullptr, state, history);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_argmax(const float *x, int64_t n, int32_t *out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && out && n > 0, "bad argmax args");
    static Cand *scratch = nullptr; // small, lives for the process
    if (!scratch) GGB_CUDA(cudaMalloc(&scratch, sizeof(Cand) * kArgmaxParts));
    argmax_stage1<<<kArgmaxParts, 256, 0, ggb_cuda_stream()>>>(x, n, scratch);
    GGB_LAUNCH_CHECK();
    argmax_stage2<<<1, 256, 0, ggb_cuda_stream()>>>(scratch, kArgmaxParts, out, nullptr, nullptr);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

} // extern "C"
