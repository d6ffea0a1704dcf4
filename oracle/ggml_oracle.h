/*
 * ggml_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the ggml numerics that gemma.ggml's hot path runs
 * through (the reference links an un-vendored, patched ggml under lib/, absent
 * from /root/reference: CMakeLists.txt:13, .gitignore:4).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libggb_cuda.so / libggb_host.so) never
 * links or calls it.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own fixtures — the
 * reference commits no golden tensors.  What pins it instead:
 *   - Q4_0 / Q8_0 codecs: known-answer vectors from gguf-py 0.19.0
 *     (tests/golden/, generator tests/golden/make_golden.py), whose Q8_0 is
 *     documented bit-exact to ggml-quants.c;
 *   - vec_dot_f16 operand types / double accumulation: src/kernals.cl:158-164;
 *   - loop / indexing / row partition: src/hpc.cpp:15-41, 245-273;
 *   - graph order and shapes: src/gemma_model.cpp:438-530, 665-747.
 *
 * There is ONE accumulation order: the scalar ggml order (one accumulator, elements / blocks in
 * order; double accumulators where ggml's scalar code has them).  The sm_100a kernels reproduce it
 * bit for bit, so every comparison in tests/ is a comparison of bits.  (`order` parameters are kept
 * in a few signatures for ABI stability and are ignored.)
 * Compile with -ffp-contract=off: every fp32 expression below is evaluated
 * with one IEEE rounding per operator, left to right, as ISO C specifies.
 */
#ifndef GGML_ORACLE_H
#define GGML_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_QK 32 /* QK4_0 == QK8_0 == 32 */

/* ggml_type ids (upstream enum values; src/hpc.cpp:132-143 switches on them) */
enum orc_type { ORC_TYPE_F32 = 0, ORC_TYPE_F16 = 1, ORC_TYPE_Q4_0 = 2, ORC_TYPE_Q8_0 = 8 };

#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[ORC_QK / 2]; } orc_block_q4_0; /* 18 B */
typedef struct { uint16_t d; int8_t qs[ORC_QK]; } orc_block_q8_0;      /* 34 B */
#pragma pack(pop)


/* ---- fp16 ---- */
float orc_fp16_to_fp32(uint16_t h);
uint16_t orc_fp32_to_fp16(float f); /* round-to-nearest-even */

/* ---- codecs ---- */
size_t orc_row_size(int type, int64_t k);
void orc_quantize_row_q4_0(const float *x, orc_block_q4_0 *y, int64_t k);
void orc_quantize_row_q8_0(const float *x, orc_block_q8_0 *y, int64_t k);
void orc_dequantize_row_q4_0(const orc_block_q4_0 *x, float *y, int64_t k);
void orc_dequantize_row_q8_0(const orc_block_q8_0 *x, float *y, int64_t k);
void orc_dequantize_row(int type, const void *x, float *y, int64_t k);

/* ---- dot products ---- */
/* scalar ggml order (blocks in order, one fp32 accumulator) */
void orc_vec_dot_q4_0_q8_0(int n, float *s, const void *vx, const void *vy);
void orc_vec_dot_q8_0_q8_0(int n, float *s, const void *vx, const void *vy);
/* fp16·fp16, accumulated in double (src/kernals.cl:158-164). */
void orc_vec_dot_f16(int n, float *s, const uint16_t *x, const uint16_t *y);
/* generic dispatcher (order: ignored) */
void orc_vec_dot(int type, int order, int n, float *s, const void *vx, const void *vy);

/* ---- mul_mat as the reference's hpc.cpp runs it (single thread) ---- */
void orc_mul_mat_sub(int64_t start_row, int64_t end_row, int64_t shared_edge, int64_t col_num,
                     int64_t ne01, int64_t ne1, int64_t nb01, int64_t nb1, int64_t nb2,
                     int64_t row_size, const char *src0, const char *src1, char *dst,
                     int type, int order);

/* ---- element-wise / normalisation ops ---- */
void orc_rms_norm(int64_t n, const float *x, float eps, float *y);     /* REF: sequential double sum */
void orc_mul(int64_t n, const float *a, const float *b, float *y);
void orc_add(int64_t n, const float *a, const float *b, float *y);
void orc_scale(int64_t n, const float *x, float s, float *y);
/* rope NEOX (mode 2): cos/sin recurrence for one position, n_dims/2 entries each */
void orc_rope_cache(int32_t pos, int n_dims, float freq_base, float *cos_out, float *sin_out);
void orc_rope_neox(int n_dims, int n_heads, int32_t pos, float freq_base, const float *x, float *y);
/* fp16 lookup tables (64 Ki entries each) */
void orc_build_exp_table(uint16_t *table);
void orc_build_gelu_table(uint16_t *table);
const uint16_t *orc_exp_table(void);
const uint16_t *orc_gelu_table(void);
void orc_gelu(int64_t n, const float *x, float *y);
/* soft_max_ext with max_bias = 0, pos = NULL; mask may be NULL */
void orc_soft_max_ext(int64_t nc, const float *x, const float *mask, float scale, float *y);
/* first maximum, strict > (src/gemma_model.cpp:532-546) */
int32_t orc_argmax(int64_t n, const float *x);

/* ---- decode attention for one token ----
 * q: f32 [H][d] (already roped and scaled); kcache/vcache: fp16, position-major
 * [n_ctx][Hkv*d]; n_pos = number of valid positions (pos+1).  out f32 [H][d].
 * REF restates kq = K16·f16(q) (double acc) -> soft_max_ext -> f16(P)·V16 (double acc),
 * over n_kv = roundup32-padded length with masked tail (src/gemma_model.cpp:326-335,428-436).
 */
void orc_attn_decode_ref(int H, int Hkv, int d, int64_t n_pos, int64_t n_ctx, const float *q,
                         const uint16_t *kcache, const uint16_t *vcache, int64_t kv_stride, float *out);

/* ---- synthetic weights: counter-based, integer-exact ---- */
uint64_t orc_splitmix64(uint64_t x);
float orc_synth_value(uint64_t seed, uint64_t tensor_id, uint64_t index, float sigma);
void orc_synth_row(uint64_t seed, uint64_t tensor_id, int64_t row, int64_t k, float sigma, float bias, float *out);


/* ---- K-quants: the formats the reference's own device kernels target ----
 * block layouts: src/kernals.cl:13-34 (= ggml block_q4_K 144 B, block_q6_K 210 B, block_q8_K 292 B; QK_K = 256) */
#define ORC_QK_K 256
#define ORC_TYPE_Q4_K 12
#define ORC_TYPE_Q6_K 14
#define ORC_TYPE_Q8_K 15
#pragma pack(push, 1)
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qs[128]; } orc_block_q4_K;
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; uint16_t d; } orc_block_q6_K;
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; } orc_block_q8_K;
#pragma pack(pop)
/* [UPSTREAM] quantize_row_q8_K_reference: the activation quantiser of the K-quant mat-muls (ggml mul_mat INIT) */
void orc_quantize_row_q8_K(const float *x, orc_block_q8_K *y, int64_t k);
/* ref: src/kernals.cl:48-111 (restated; pinned against oracle/_ref) */
void orc_vec_dot_q4_K_q8_K(int n, float *s, const void *vx, const void *vy);
/* [UPSTREAM] ggml_vec_dot_q6_K_q8_K, scalar branch, QK_K = 256.  (The reference's copy, src/kernals.cl:114-156, is ggml's
 * QK_K == 64 branch under QK_K = 256: it unpacks 64 of the 256 weights and reads the rest uninitialised - not followed.) */
void orc_vec_dot_q6_K_q8_K(int n, float *s, const void *vx, const void *vy);
/* [UPSTREAM] dequantize_row_q4_K / q6_K (checked against gguf-py) */
void orc_dequantize_row_q4_K(const orc_block_q4_K *x, float *y, int64_t k);
void orc_dequantize_row_q6_K(const orc_block_q6_K *x, float *y, int64_t k);

#ifdef __cplusplus
}
#endif
#endif
