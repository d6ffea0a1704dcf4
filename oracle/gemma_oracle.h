/*
 * gemma_oracle.h — CPU ORACLE, model level (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Restates the reference's forward pass (ref: src/gemma_model.cpp:231-286, 665-747) and its
 * CPU mul_mat path (ref: src/hpc.cpp:216-277 + src/thread_pool.{h,cpp}) on synthetic
 * Gemma-architecture weights.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 */
#ifndef GEMMA_ORACLE_H
#define GEMMA_ORACLE_H

#include "ggml_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t n_embd, n_layer, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx;
    float eps, rope_base;
    int32_t wtype; /* ORC_TYPE_Q4_0 | ORC_TYPE_Q8_0, used for every 2-D matrix */
    int32_t pad_;
    uint64_t seed;
    /* synthetic-weight std: token_embd / every other matrix.  SURVEY.md §8d uses 0.02 for both;
     * tests also use a small sigma_embd so the (tied) output projection does not simply echo
     * the input token and greedy parity is a sensitive check. */
    float sigma_embd, sigma_mat;
} orc_config;

/* tensor ids, in the order the reference binds them (ref: src/gemma_model.cpp:159-194) */
enum { ORC_T_TOKEN_EMBD = 0, ORC_T_OUTPUT_NORM = 1, ORC_T_LAYER0 = 2, ORC_T_PER_LAYER = 9 };
enum { ORC_L_ATTN_NORM = 0, ORC_L_ATTN_Q, ORC_L_ATTN_K, ORC_L_ATTN_V, ORC_L_ATTN_OUT,
       ORC_L_FFN_NORM, ORC_L_FFN_GATE, ORC_L_FFN_UP, ORC_L_FFN_DOWN };

#define ORC_SYNTH_SIGMA 0.02f /* norm weights: 1 + N(0, 0.02^2) */

typedef struct orc_model orc_model;

orc_model *orc_model_create(const orc_config *cfg, int n_threads);
void orc_model_free(orc_model *m);
/* raw tensor bytes: ggml AoS blocks for matrices, f32 for norm weights */
const void *orc_model_tensor(const orc_model *m, int tensor_id, int64_t *rows, int64_t *cols, int64_t *nbytes);
void orc_model_reset(orc_model *m); /* zero the KV cache */
void orc_model_fill_kv_synth(orc_model *m, uint64_t seed);
/* Run n_tokens tokens at positions pos0.. (prefill if n_tokens > 1).  order: ignored (there is one order, the scalar ggml one).
 * logits_all != 0 computes the output projection for every token as the reference does
 * (ref: src/gemma_model.cpp:740); logits_last receives the last token's n_vocab logits. */
int orc_model_forward(orc_model *m, const int32_t *tokens, int n_tokens, int pos0, int order,
                      int logits_all, float *logits_last);
/* greedy loop (ref: src/gemma_model.cpp:548-575): prefill prompt, then n_decode tokens. */
int orc_model_generate(orc_model *m, const int32_t *prompt, int n_prompt, int n_decode, int order,
                       int32_t *out_tokens, double *prefill_s, double *decode_s);
/* captures from the last forward (last token): names "l_out", "kqv_out", "sa_out", "ffn_act"
 * (per layer) and "result_norm", "result_output" (layer ignored). */
const float *orc_model_capture(const orc_model *m, const char *name, int layer, int64_t *n);
/* KV cache rows (fp16, position-major [n_ctx][n_head_kv*head_dim]) */
uint16_t *orc_model_kcache(orc_model *m, int layer);
uint16_t *orc_model_vcache(orc_model *m, int layer);

/* the reference's mul_mat dispatch with CPU_SPLIT_RATIO 1.0 on the model's pool
 * (ref: src/hpc.cpp:216-277).  src0 = ggml AoS rows; wdata = Q8_0 rows of src1. */
void orc_hpc_mul_mat(orc_model *m, int64_t ne01, int64_t ne11, int64_t ne12, int64_t nb01, int64_t ne1,
                     int64_t nb1, int64_t nb2, size_t row_size, int64_t shared_edge, const void *src0,
                     void *dst, int src0_type, const char *wdata, int order);
int orc_model_threads(const orc_model *m);

#ifdef __cplusplus
}
#endif
#endif
