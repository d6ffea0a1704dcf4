/*
 * ggml_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See ggml_oracle.h.
 *
 * Each function cites what it follows.  "[UPSTREAM]" = semantics of the un-vendored ggml
 * (ggml.c / ggml-quants.c, llama.cpp circa Feb-Mar 2024) restated from its published
 * algorithm; "ref:" = a file:line under /root/reference.
 * Build: gcc -O3 -ffp-contract=off (no FMA contraction, no fast-math).
 *
 * PINNING.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *   - Q4_K x Q8_K and F16 x F16 dots: bit-identical to the REFERENCE'S OWN CODE (src/kernals.cl compiled where it lies
 *     into oracle/_ref, tests/test_kquants_cpu.py);
 *   - Q4_0 / Q8_0 codecs, fp16 conversion, Q4_K / Q6_K block layouts: gguf-py known-answer vectors (tests/golden/);
 *   - Q4_0 / Q8_0 dots, rms_norm, soft_max_ext, rope, GeGLU, the forward pass: the reference commits no fixture and cannot
 *     be built here (its patched ggml is un-vendored), so these are PINNED TO THE PUBLISHED SCALAR ggml ALGORITHM AND TO
 *     float64 MODELS, NOT TO REFERENCE OUTPUT ("parity unpinned" in the tier's sense for exactly these functions).
 */
#include "ggml_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ fp16 */

/* [UPSTREAM] GGML_FP16_TO_FP32: IEEE binary16 -> binary32, exact. */
float orc_fp16_to_fp32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do {
                man <<= 1;
                e++;
            } while ((man & 0x400u) == 0);
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* [UPSTREAM] GGML_FP32_TO_FP16: binary32 -> binary16, round-to-nearest-even (F16C semantics). */
uint16_t orc_fp32_to_fp16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    const uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) { /* inf / nan */
        if (absx > 0x7f800000u) return (uint16_t)(sign | 0x7e00u | ((absx >> 13) & 0x3ffu));
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx >= 0x477ff000u) { /* >= 65520 rounds to inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x33000001u) { /* <= 2^-25 rounds to zero (tie at exactly 2^-25 -> even = 0) */
        return sign;
    }
    int32_t e = (int32_t)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u; /* 24-bit significand */
    uint32_t shift, half_bits;
    if (e < -14) { /* subnormal half: value = m * 2^(e-23); unit = 2^-24 */
        shift = (uint32_t)(-1 - e); /* m >> shift gives multiples of 2^-24: 23 - (e + 24) */
        half_bits = 0;
    } else {
        shift = 13;
        half_bits = (uint32_t)(e + 15) << 10;
        m &= 0x7fffffu;
    }
    uint32_t q = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u);
    const uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) q++;
    /* carry out of the mantissa naturally bumps the exponent */
    return (uint16_t)(sign | (half_bits + q));
}

/* ---------------------------------------------------------------- codecs */

size_t orc_row_size(int type, int64_t k) {
    switch (type) {
    case ORC_TYPE_F32: return (size_t)k * 4;
    case ORC_TYPE_F16: return (size_t)k * 2;
    case ORC_TYPE_Q4_0: return (size_t)(k / ORC_QK) * sizeof(orc_block_q4_0);
    case ORC_TYPE_Q8_0: return (size_t)(k / ORC_QK) * sizeof(orc_block_q8_0);
    case ORC_TYPE_Q4_K: return (size_t)(k / ORC_QK_K) * sizeof(orc_block_q4_K);
    case ORC_TYPE_Q6_K: return (size_t)(k / ORC_QK_K) * sizeof(orc_block_q6_K);
    case ORC_TYPE_Q8_K: return (size_t)(k / ORC_QK_K) * sizeof(orc_block_q8_K);
    default: return 0;
    }
}

/* [UPSTREAM] quantize_row_q4_0_reference; cross-checked vs gguf/quants.py:222-238. */
void orc_quantize_row_q4_0(const float *x, orc_block_q4_0 *y, int64_t k) {
    const int64_t nb = k / ORC_QK;
    for (int64_t i = 0; i < nb; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < ORC_QK; j++) {
            const float v = x[i * ORC_QK + j];
            if (amax < fabsf(v)) {
                amax = fabsf(v);
                max = v;
            }
        }
        const float d = max / -8;
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = orc_fp32_to_fp16(d);
        for (int j = 0; j < ORC_QK / 2; ++j) {
            const float x0 = x[i * ORC_QK + 0 + j] * id;
            const float x1 = x[i * ORC_QK + ORC_QK / 2 + j] * id;
            int a = (int)(int8_t)(x0 + 8.5f);
            int b = (int)(int8_t)(x1 + 8.5f);
            const uint8_t xi0 = (uint8_t)(a < 15 ? a : 15);
            const uint8_t xi1 = (uint8_t)(b < 15 ? b : 15);
            y[i].qs[j] = (uint8_t)(xi0 | (xi1 << 4));
        }
    }
}

/* [UPSTREAM] quantize_row_q8_0_reference (scalar: roundf, id = 1/d); gguf/quants.py:381-393. */
void orc_quantize_row_q8_0(const float *x, orc_block_q8_0 *y, int64_t k) {
    const int64_t nb = k / ORC_QK;
    for (int64_t i = 0; i < nb; i++) {
        float amax = 0.0f;
        for (int j = 0; j < ORC_QK; j++) {
            const float v = fabsf(x[i * ORC_QK + j]);
            if (amax < v) amax = v;
        }
        const float d = amax / ((1 << 7) - 1);
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = orc_fp32_to_fp16(d);
        for (int j = 0; j < ORC_QK; ++j) {
            const float x0 = x[i * ORC_QK + j] * id;
            y[i].qs[j] = (int8_t)roundf(x0);
        }
    }
}

/* [UPSTREAM] dequantize_row_q4_0: (q - 8) * d; element j = low nibble, j+16 = high nibble. */
void orc_dequantize_row_q4_0(const orc_block_q4_0 *x, float *y, int64_t k) {
    const int64_t nb = k / ORC_QK;
    for (int64_t i = 0; i < nb; i++) {
        const float d = orc_fp16_to_fp32(x[i].d);
        for (int j = 0; j < ORC_QK / 2; ++j) {
            const int x0 = (x[i].qs[j] & 0x0F) - 8;
            const int x1 = (x[i].qs[j] >> 4) - 8;
            y[i * ORC_QK + j] = x0 * d;
            y[i * ORC_QK + j + ORC_QK / 2] = x1 * d;
        }
    }
}

/* [UPSTREAM] dequantize_row_q8_0: q * d. */
void orc_dequantize_row_q8_0(const orc_block_q8_0 *x, float *y, int64_t k) {
    const int64_t nb = k / ORC_QK;
    for (int64_t i = 0; i < nb; i++) {
        const float d = orc_fp16_to_fp32(x[i].d);
        for (int j = 0; j < ORC_QK; ++j) y[i * ORC_QK + j] = x[i].qs[j] * d;
    }
}

void orc_dequantize_row(int type, const void *x, float *y, int64_t k) {
    if (type == ORC_TYPE_Q4_0) orc_dequantize_row_q4_0((const orc_block_q4_0 *)x, y, k);
    else if (type == ORC_TYPE_Q8_0) orc_dequantize_row_q8_0((const orc_block_q8_0 *)x, y, k);
    else if (type == ORC_TYPE_Q4_K) orc_dequantize_row_q4_K((const orc_block_q4_K *)x, y, k);
    else if (type == ORC_TYPE_Q6_K) orc_dequantize_row_q6_K((const orc_block_q6_K *)x, y, k);
    else if (type == ORC_TYPE_F32) memcpy(y, x, (size_t)k * 4);
    else if (type == ORC_TYPE_F16) {
        for (int64_t i = 0; i < k; i++) y[i] = orc_fp16_to_fp32(((const uint16_t *)x)[i]);
    }
}

/* ---------------------------------------------------------- dot products */

static inline int q4_block_sumi(const orc_block_q4_0 *x, const orc_block_q8_0 *y) {
    int sumi = 0;
    for (int j = 0; j < ORC_QK / 2; ++j) {
        const int v0 = (x->qs[j] & 0x0F) - 8;
        const int v1 = (x->qs[j] >> 4) - 8;
        sumi += (v0 * y->qs[j]) + (v1 * y->qs[j + ORC_QK / 2]);
    }
    return sumi;
}

static inline int q8_block_sumi(const orc_block_q8_0 *x, const orc_block_q8_0 *y) {
    int sumi = 0;
    for (int j = 0; j < ORC_QK; j++) sumi += x->qs[j] * y->qs[j];
    return sumi;
}

/* one block's contribution, exactly as the scalar loops write it:
 * q4_0: sumi*d_x*d_y  (left to right)      q8_0: sumi*(d_x*d_y) */
static inline float q4_block_term(const orc_block_q4_0 *x, const orc_block_q8_0 *y) {
    return q4_block_sumi(x, y) * orc_fp16_to_fp32(x->d) * orc_fp16_to_fp32(y->d);
}
static inline float q8_block_term(const orc_block_q8_0 *x, const orc_block_q8_0 *y) {
    return q8_block_sumi(x, y) * (orc_fp16_to_fp32(x->d) * orc_fp16_to_fp32(y->d));
}

/* [UPSTREAM] ggml_vec_dot_q4_0_q8_0, scalar branch; called through the fn pointer at
 * ref: src/hpc.cpp:35 with nrc = 1. */
void orc_vec_dot_q4_0_q8_0(int n, float *s, const void *vx, const void *vy) {
    const int nb = n / ORC_QK;
    const orc_block_q4_0 *x = (const orc_block_q4_0 *)vx;
    const orc_block_q8_0 *y = (const orc_block_q8_0 *)vy;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) sumf += q4_block_term(&x[i], &y[i]);
    *s = sumf;
}

/* [UPSTREAM] ggml_vec_dot_q8_0_q8_0, scalar branch. */
void orc_vec_dot_q8_0_q8_0(int n, float *s, const void *vx, const void *vy) {
    const int nb = n / ORC_QK;
    const orc_block_q8_0 *x = (const orc_block_q8_0 *)vx;
    const orc_block_q8_0 *y = (const orc_block_q8_0 *)vy;
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) sumf += q8_block_term(&x[i], &y[i]);
    *s = sumf;
}


/* [UPSTREAM] ggml_vec_dot_f16 scalar branch; ref: src/kernals.cl:158-164 (double accumulator). */
void orc_vec_dot_f16(int n, float *s, const uint16_t *x, const uint16_t *y) {
    double sumf = 0.0;
    for (int i = 0; i < n; ++i) sumf += (double)(orc_fp16_to_fp32(x[i]) * orc_fp16_to_fp32(y[i]));
    *s = (float)sumf;
}

void orc_vec_dot(int type, int order, int n, float *s, const void *vx, const void *vy) {
    (void)order; /* there is one order: the scalar ggml one */
    if (type == ORC_TYPE_Q4_0) {
        orc_vec_dot_q4_0_q8_0(n, s, vx, vy);
    } else if (type == ORC_TYPE_Q8_0) {
        orc_vec_dot_q8_0_q8_0(n, s, vx, vy);
    } else if (type == ORC_TYPE_F16) {
        orc_vec_dot_f16(n, s, (const uint16_t *)vx, (const uint16_t *)vy);
    } else if (type == ORC_TYPE_Q4_K) {
        orc_vec_dot_q4_K_q8_K(n, s, vx, vy);
    } else if (type == ORC_TYPE_Q6_K) {
        orc_vec_dot_q6_K_q8_K(n, s, vx, vy);
    } else {
        *s = NAN;
    }
}

/* ref: src/hpc.cpp:15-41 (mul_mat_sub with MUL_MAT_BLOCK_SIZE 1, src/macro.h:23):
 * dst[(col % ne1)*nb1 + (col / ne1)*nb2 + row*4] = vec_dot(K, src0 + row*nb01, src1 + col*row_size). */
void orc_mul_mat_sub(int64_t start_row, int64_t end_row, int64_t shared_edge, int64_t col_num,
                     int64_t ne01, int64_t ne1, int64_t nb01, int64_t nb1, int64_t nb2,
                     int64_t row_size, const char *src0, const char *src1, char *dst,
                     int type, int order) {
    for (int64_t row0 = start_row; row0 < end_row && row0 < ne01; row0++) {
        for (int64_t col_i = 0; col_i < col_num; col_i++) {
            const int64_t mat_i = col_i / ne1;
            const int64_t mat_col_i = col_i % ne1;
            const char *src1_col = src1 + col_i * row_size;
            float *dst_col = (float *)(dst + (mat_col_i * nb1 + mat_i * nb2));
            orc_vec_dot(type, order, (int)shared_edge, &dst_col[row0], src0 + row0 * nb01, src1_col);
        }
    }
}

/* ------------------------------------------------------------ element-wise */

/* [UPSTREAM] ggml_compute_forward_rms_norm_f32: sum in double of (double)(x*x) with the product
 * rounded to fp32 first; mean = sum/n (to float); scale = 1/sqrtf(mean + eps). */
void orc_rms_norm(int64_t n, const float *x, float eps, float *y) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; i++) sum += (double)(x[i] * x[i]);
    const float mean = (float)(sum / (double)n);
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = 0; i < n; i++) y[i] = x[i] * scale;
}


void orc_mul(int64_t n, const float *a, const float *b, float *y) {
    for (int64_t i = 0; i < n; i++) y[i] = a[i] * b[i];
}
void orc_add(int64_t n, const float *a, const float *b, float *y) {
    for (int64_t i = 0; i < n; i++) y[i] = a[i] + b[i];
}
void orc_scale(int64_t n, const float *x, float s, float *y) {
    for (int64_t i = 0; i < n; i++) y[i] = x[i] * s;
}

/* [UPSTREAM] ggml_compute_forward_rope_f32, NEOX branch (mode & 2), ext_factor = 0,
 * freq_scale = 1, attn_factor = 1 (ref: src/macro.h:12-18, src/gemma_model.cpp:698-716):
 * theta_scale = powf(freq_base, -2/n_dims); theta starts at pos and is multiplied each step. */
void orc_rope_cache(int32_t pos, int n_dims, float freq_base, float *cos_out, float *sin_out) {
    const float theta_scale = powf(freq_base, -2.0f / n_dims);
    float theta = (float)pos;
    for (int i = 0; i < n_dims / 2; i++) {
        cos_out[i] = cosf(theta);
        sin_out[i] = sinf(theta);
        theta *= theta_scale;
    }
}

/* x, y: [n_heads][n_dims]; pairs (i, i + n_dims/2). */
void orc_rope_neox(int n_dims, int n_heads, int32_t pos, float freq_base, const float *x, float *y) {
    float *c = (float *)malloc(sizeof(float) * (size_t)n_dims);
    float *s = c + n_dims / 2;
    orc_rope_cache(pos, n_dims, freq_base, c, s);
    for (int h = 0; h < n_heads; h++) {
        const float *xh = x + (size_t)h * n_dims;
        float *yh = y + (size_t)h * n_dims;
        for (int i = 0; i < n_dims / 2; i++) {
            const float x0 = xh[i];
            const float x1 = xh[i + n_dims / 2];
            yh[i] = x0 * c[i] - x1 * s[i];
            yh[i + n_dims / 2] = x0 * s[i] + x1 * c[i];
        }
    }
    free(c);
}

/* [UPSTREAM] ggml_init table construction: table[i] = FP32_TO_FP16(f(FP16_TO_FP32(i))). */
void orc_build_exp_table(uint16_t *table) {
    for (uint32_t i = 0; i < 65536; i++) table[i] = orc_fp32_to_fp16(expf(orc_fp16_to_fp32((uint16_t)i)));
}
static inline float gelu_f32(float x) {
    const float GELU_COEF_A = 0.044715f;
    const float SQRT_2_OVER_PI = 0.79788456080286535587989211986876f;
    return 0.5f * x * (1.0f + tanhf(SQRT_2_OVER_PI * x * (1.0f + GELU_COEF_A * x * x)));
}
void orc_build_gelu_table(uint16_t *table) {
    for (uint32_t i = 0; i < 65536; i++) table[i] = orc_fp32_to_fp16(gelu_f32(orc_fp16_to_fp32((uint16_t)i)));
}

static uint16_t g_exp_table[65536], g_gelu_table[65536];
static int g_tables_ready = 0;
static void ensure_tables(void) {
    if (!g_tables_ready) {
        orc_build_exp_table(g_exp_table);
        orc_build_gelu_table(g_gelu_table);
        g_tables_ready = 1;
    }
}
const uint16_t *orc_exp_table(void) {
    ensure_tables();
    return g_exp_table;
}
const uint16_t *orc_gelu_table(void) {
    ensure_tables();
    return g_gelu_table;
}

/* [UPSTREAM] ggml_vec_gelu_f32 with GGML_GELU_FP16: y = table[f16(x)]. */
void orc_gelu(int64_t n, const float *x, float *y) {
    const uint16_t *t = orc_gelu_table();
    for (int64_t i = 0; i < n; i++) y[i] = orc_fp16_to_fp32(t[orc_fp32_to_fp16(x[i])]);
}

/* [UPSTREAM] ggml_compute_forward_soft_max_f32 (max_bias = 0, no pos): w = x*scale + mask;
 * p = exp_table[f16(w - max)] (0 where w == -inf); sum in double; p *= (float)(1/sum). */
void orc_soft_max_ext(int64_t nc, const float *x, const float *mask, float scale, float *y) {
    const uint16_t *t = orc_exp_table();
    float *wp = (float *)malloc(sizeof(float) * (size_t)nc);
    for (int64_t i = 0; i < nc; i++) wp[i] = x[i];
    for (int64_t i = 0; i < nc; i++) wp[i] *= scale;
    if (mask)
        for (int64_t i = 0; i < nc; i++) wp[i] += mask[i];
    float max = -INFINITY;
    for (int64_t i = 0; i < nc; i++) max = wp[i] > max ? wp[i] : max;
    double sum = 0.0;
    for (int64_t i = 0; i < nc; i++) {
        if (wp[i] == -INFINITY) {
            y[i] = 0.0f;
        } else {
            const float val = orc_fp16_to_fp32(t[orc_fp32_to_fp16(wp[i] - max)]);
            sum += (double)val;
            y[i] = val;
        }
    }
    sum = 1.0 / sum;
    const float fs = (float)sum;
    for (int64_t i = 0; i < nc; i++) y[i] *= fs;
    free(wp);
}

/* ref: src/gemma_model.cpp:532-546 */
int32_t orc_argmax(int64_t n, const float *x) {
    float max_val = -INFINITY;
    int32_t max_idx = -1;
    for (int64_t i = 0; i < n; i++) {
        if (x[i] > max_val) {
            max_val = x[i];
            max_idx = (int32_t)i;
        }
    }
    return max_idx;
}

/* -------------------------------------------------------------- attention */

/* REF decode attention for one token (ref: src/gemma_model.cpp:454-496 graph_build_kqv;
 * mask / n_kv: :326-335, :428-436).  q is converted to fp16 by the mul_mat INIT stage
 * (vec_dot_type of F16 is F16), as are the soft-max probabilities before kqv. */
void orc_attn_decode_ref(int H, int Hkv, int d, int64_t n_pos, int64_t n_ctx, const float *q,
                         const uint16_t *kcache, const uint16_t *vcache, int64_t kv_stride, float *out) {
    int64_t n_kv = ((n_pos / 32) + 1) * 32; /* input.size() == n_pos for decode */
    if (n_kv > n_ctx) n_kv = n_ctx;
    const int G = H / Hkv;
    uint16_t *q16 = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)d);
    float *kq = (float *)malloc(sizeof(float) * (size_t)n_kv);
    float *mask = (float *)malloc(sizeof(float) * (size_t)n_kv);
    float *p = (float *)malloc(sizeof(float) * (size_t)n_kv);
    uint16_t *p16 = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n_kv);
    uint16_t *vcol = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n_kv);
    for (int64_t j = 0; j < n_kv; j++) mask[j] = (j > n_pos - 1) ? -INFINITY : 0.0f;
    for (int h = 0; h < H; h++) {
        const int g = h / G;
        for (int i = 0; i < d; i++) q16[i] = orc_fp32_to_fp16(q[(size_t)h * d + i]);
        for (int64_t pz = 0; pz < n_kv; pz++)
            orc_vec_dot_f16(d, &kq[pz], kcache + pz * kv_stride + (size_t)g * d, q16);
        orc_soft_max_ext(n_kv, kq, mask, 1.0f, p);
        for (int64_t pz = 0; pz < n_kv; pz++) p16[pz] = orc_fp32_to_fp16(p[pz]);
        for (int i = 0; i < d; i++) {
            for (int64_t pz = 0; pz < n_kv; pz++) vcol[pz] = vcache[pz * kv_stride + (size_t)g * d + i];
            orc_vec_dot_f16((int)n_kv, &out[(size_t)h * d + i], vcol, p16);
        }
    }
    free(q16); free(kq); free(mask); free(p); free(p16); free(vcol);
}

/* ------------------------------------------------------- synthetic weights */

uint64_t orc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* Approximately N(0, sigma^2): Irwin-Hall sum of four 16-bit uniforms, centred, scaled by ONE
 * fp32 multiply of an exactly representable integer -> bit-reproducible on CPU and GPU
 * (no libm).  std of the integer sum = sqrt(4*(65536^2-1)/12) = 37837.2274... */
float orc_synth_value(uint64_t seed, uint64_t tensor_id, uint64_t index, float sigma) {
    const uint64_t r = orc_splitmix64(orc_splitmix64(seed + tensor_id * 0xD1B54A32D192ED03ull) + index);
    const int32_t sum = (int32_t)(r & 0xffff) + (int32_t)((r >> 16) & 0xffff) + (int32_t)((r >> 32) & 0xffff) +
                        (int32_t)((r >> 48) & 0xffff) - 131070;
    const float unit = sigma * (1.0f / 37837.2274f);
    return (float)sum * unit;
}

void orc_synth_row(uint64_t seed, uint64_t tensor_id, int64_t row, int64_t k, float sigma, float bias, float *out) {
    for (int64_t i = 0; i < k; i++)
        out[i] = bias + orc_synth_value(seed, tensor_id, (uint64_t)(row * k + i), sigma);
}

/* ------------------------------------------------------- K-quants */

static inline int orc_nearest_int(float fval) { /* [UPSTREAM] ggml-quants.c nearest_int: round half to even */
    float val = fval + 12582912.f;
    int i;
    memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

/* [UPSTREAM] quantize_row_q8_K_reference */
void orc_quantize_row_q8_K(const float *x, orc_block_q8_K *y, int64_t k) {
    const int64_t nb = k / ORC_QK_K;
    for (int64_t i = 0; i < nb; i++) {
        float max = 0, amax = 0;
        for (int j = 0; j < ORC_QK_K; ++j) {
            const float ax = fabsf(x[j]);
            if (ax > amax) { amax = ax; max = x[j]; }
        }
        if (!amax) {
            y[i].d = 0;
            memset(y[i].qs, 0, ORC_QK_K);
            memset(y[i].bsums, 0, sizeof(y[i].bsums)); /* (ggml leaves bsums untouched here; zero is what a zeroed wdata holds) */
            x += ORC_QK_K;
            continue;
        }
        const float iscale = -128.f / max;
        for (int j = 0; j < ORC_QK_K; ++j) {
            int v = orc_nearest_int(iscale * x[j]);
            y[i].qs[j] = (int8_t)(v < 127 ? v : 127);
        }
        for (int j = 0; j < ORC_QK_K / 16; ++j) {
            int sum = 0;
            for (int ii = 0; ii < 16; ++ii) sum += y[i].qs[j * 16 + ii];
            y[i].bsums[j] = (int16_t)sum;
        }
        y[i].d = 1 / iscale;
        x += ORC_QK_K;
    }
}

/* the 6-bit scales / mins of a Q4_K block, unpacked as the reference does (src/kernals.cl:75-80) */
static void q4k_scales(const uint8_t *packed, uint8_t *scales /*8*/, uint8_t *mins /*8*/) {
    const uint32_t kmask1 = 0x3f3f3f3f, kmask2 = 0x0f0f0f0f, kmask3 = 0x03030303;
    uint32_t utmp[4];
    memcpy(utmp, packed, 12);
    utmp[3] = ((utmp[2] >> 4) & kmask2) | (((utmp[1] >> 6) & kmask3) << 4);
    const uint32_t uaux = utmp[1] & kmask1;
    utmp[1] = (utmp[2] & kmask2) | (((utmp[0] >> 6) & kmask3) << 4);
    utmp[2] = uaux;
    utmp[0] &= kmask1;
    memcpy(scales, &utmp[0], 8);
    memcpy(mins, &utmp[2], 8);
}

/* ref: src/kernals.cl:48-111.  Integer part exact; fp32 part: eight lane sums (lane = element index mod 8) advanced
 * block by block with d * aux32[l], the mins term subtracted from sumf block by block, lanes added at the end. */
void orc_vec_dot_q4_K_q8_K(int n, float *s, const void *vx, const void *vy) {
    const orc_block_q4_K *x = (const orc_block_q4_K *)vx;
    const orc_block_q8_K *y = (const orc_block_q8_K *)vy;
    const int nb = n / ORC_QK_K;
    int8_t aux8[ORC_QK_K];
    float sums[8] = {0};
    float sumf = 0;
    for (int i = 0; i < nb; ++i) {
        const uint8_t *q4 = x[i].qs;
        const int8_t *q8 = y[i].qs;
        int32_t aux32[8] = {0};
        int8_t *a = aux8;
        for (int j = 0; j < ORC_QK_K / 64; ++j) {
            for (int l = 0; l < 32; ++l) a[l] = (int8_t)(q4[l] & 0xF);
            a += 32;
            for (int l = 0; l < 32; ++l) a[l] = (int8_t)(q4[l] >> 4);
            a += 32; q4 += 32;
        }
        uint8_t scales[8], mins[8];
        q4k_scales(x[i].scales, scales, mins);
        int sumi = 0;
        for (int j = 0; j < ORC_QK_K / 16; ++j) sumi += y[i].bsums[j] * mins[j / 2];
        a = aux8;
        for (int j = 0; j < ORC_QK_K / 32; ++j) {
            const int32_t scale = scales[j];
            for (int g = 0; g < 4; ++g) {
                for (int l = 0; l < 8; ++l) aux32[l] += scale * (int16_t)(q8[l] * a[l]);
                q8 += 8; a += 8;
            }
        }
        const float d = orc_fp16_to_fp32(x[i].d) * y[i].d;
        for (int l = 0; l < 8; ++l) sums[l] += d * aux32[l];
        const float dmin = orc_fp16_to_fp32(x[i].dmin) * y[i].d;
        sumf -= dmin * sumi;
    }
    for (int l = 0; l < 8; ++l) sumf += sums[l];
    *s = sumf;
}

/* a Q6_K block's 256 weights as int8 (q - 32), [UPSTREAM] ggml QK_K = 256 layout */
static void q6k_unpack(const orc_block_q6_K *b, int8_t *a) {
    const uint8_t *q4 = b->ql, *qh = b->qh;
    for (int j = 0; j < ORC_QK_K; j += 128) {
        for (int l = 0; l < 32; ++l) {
            a[l + 0] = (int8_t)((q4[l + 0] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
            a[l + 32] = (int8_t)((q4[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
            a[l + 64] = (int8_t)((q4[l + 0] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
            a[l + 96] = (int8_t)((q4[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
        }
        a += 128; q4 += 64; qh += 32;
    }
}

void orc_vec_dot_q6_K_q8_K(int n, float *s, const void *vx, const void *vy) {
    const orc_block_q6_K *x = (const orc_block_q6_K *)vx;
    const orc_block_q8_K *y = (const orc_block_q8_K *)vy;
    const int nb = n / ORC_QK_K;
    int8_t aux8[ORC_QK_K];
    float sums[8] = {0};
    float sumf = 0;
    for (int i = 0; i < nb; ++i) {
        const int8_t *q8 = y[i].qs;
        int32_t aux32[8] = {0};
        q6k_unpack(&x[i], aux8);
        const int8_t *a = aux8;
        for (int j = 0; j < ORC_QK_K / 16; ++j) {
            const int scale = x[i].scales[j];
            for (int g = 0; g < 2; ++g) {
                for (int l = 0; l < 8; ++l) aux32[l] += scale * (int16_t)(q8[l] * a[l]);
                q8 += 8; a += 8;
            }
        }
        const float d = orc_fp16_to_fp32(x[i].d) * y[i].d;
        for (int l = 0; l < 8; ++l) sums[l] += d * aux32[l];
    }
    for (int l = 0; l < 8; ++l) sumf += sums[l];
    *s = sumf;
}

void orc_dequantize_row_q4_K(const orc_block_q4_K *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / ORC_QK_K; i++) {
        const float d = orc_fp16_to_fp32(x[i].d), dmin = orc_fp16_to_fp32(x[i].dmin);
        uint8_t sc[8], mn[8];
        q4k_scales(x[i].scales, sc, mn);
        const uint8_t *q = x[i].qs;
        for (int j = 0; j < 4; j++) {
            const float d1 = d * sc[2 * j], m1 = dmin * mn[2 * j], d2 = d * sc[2 * j + 1], m2 = dmin * mn[2 * j + 1];
            for (int l = 0; l < 32; ++l) *y++ = d1 * (q[l] & 0xF) - m1;
            for (int l = 0; l < 32; ++l) *y++ = d2 * (q[l] >> 4) - m2;
            q += 32;
        }
    }
}

void orc_dequantize_row_q6_K(const orc_block_q6_K *x, float *y, int64_t k) {
    int8_t a[ORC_QK_K];
    for (int64_t i = 0; i < k / ORC_QK_K; i++) {
        const float d = orc_fp16_to_fp32(x[i].d);
        q6k_unpack(&x[i], a);
        for (int j = 0; j < ORC_QK_K; j++) y[i * ORC_QK_K + j] = d * x[i].scales[j / 16] * a[j];
    }
}
