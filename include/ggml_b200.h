/*
 * ggml_b200.h — the C-ABI CUDA layer (libggb_cuda.so) that replaces the reference's OpenCL
 * offload: src/opencl.{h,cpp} (device runtime) and src/kernals.cl (kernels), and is what
 * the rewired src/hpc.cpp dispatch and the graph builder of src/gemma_model.cpp call.
 *
 * Plain C: pointers, sizes, ints.  No torch / C++ types.  Every entry returns GGB_OK (0) or a
 * negative ggb_status; ggb_last_error() gives the text (the reference exits the process on
 * error, src/opencl.cpp:13-17 — a library cannot, so failures are returned and are loud).
 * There is no CPU fallback anywhere below this header: with no CUDA device ggb_init fails.
 *
 * Unless stated otherwise "device pointer" arguments come from ggb_buffer_alloc and every
 * ggb_op_* call is ASYNCHRONOUS on the library's stream; ggb_sync() waits
 * (ref: wait_queue_finish, src/opencl.h:37).
 */
#ifndef GGML_B200_H
#define GGML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GGB_OK = 0,
    GGB_ERR_CUDA = -1,        /* a CUDA runtime call failed */
    GGB_ERR_ARG = -2,         /* bad argument / unsupported shape */
    GGB_ERR_NOT_INIT = -3,    /* ggb_init not called or failed */
    GGB_ERR_UNSUPPORTED = -4, /* tensor type without a kernel (ref: "kernel is null", src/hpc.cpp:162-166) */
    GGB_ERR_NCCL = -5
} ggb_status;

/* ggml_type values the path uses (upstream enum ids; ref: src/hpc.cpp:132-143) */
enum { GGB_TYPE_F32 = 0, GGB_TYPE_F16 = 1, GGB_TYPE_Q4_0 = 2, GGB_TYPE_Q8_0 = 8,
       GGB_TYPE_Q4_K = 12, GGB_TYPE_Q6_K = 14, GGB_TYPE_Q8_K = 15 }; /* ggml_type ids */

/* Arithmetic: every result of the Q4_0 / Q8_0 decode path and of the K-quant operators is BIT-IDENTICAL to the scalar
 * ggml code the reference's CPU path runs (one fp32 accumulator per row, blocks in order; double accumulators where
 * ggml has them) - there is no kernel-specific summation order.  The constants below are tiling
 * parameters only: blocks per K segment of the decode kernel's pass stream, threads of the rms_norm partial sums
 * (double: insensitive to their grouping at fp32 precision), positions per attention item. */
#define GGB_DEV_SEG_BLOCKS 64
#define GGB_DEV_NORM_THREADS 256
#define GGB_DEV_ATTN_CHUNK 32

/* ------------------------------------------------------------------ runtime
 * replaces: init_opencl / set_platform_and_device / setup_opencl_env / build_all_kernels
 * (src/opencl.h:22-30, src/opencl.cpp:161-245, 350-356), print_now_using (:100),
 * wait_queue_finish (:309), release_opencl_env (:247). */
int ggb_init(int device);
int ggb_shutdown(void);
const char *ggb_last_error(void);
int ggb_device_info(char *buf, size_t buflen); /* name, SMs, memory, compute capability */
int ggb_sm_count(void);
int ggb_sync(void);
void *ggb_stream(void); /* the cudaStream_t all ops are enqueued on */
int64_t ggb_launch_count(void); /* kernels launched (or replayed via graphs) since init */

/* replaces: create_buffer / release_buffer_in_set (src/opencl.h:35,38; src/opencl.cpp:313-348).
 * Buffers are device-resident and live until freed (the reference re-wraps host memory per
 * mul_mat call, src/hpc.cpp:168-170,196 — only viable on a zero-copy iGPU). */
int ggb_buffer_alloc(size_t bytes, void **dptr);
int ggb_buffer_free(void *dptr);
int ggb_buffer_memset(void *dptr, int value, size_t bytes);
int ggb_upload(void *dst_dev, const void *src_host, size_t bytes);   /* async if src is pinned */
int ggb_download(void *dst_host, const void *src_dev, size_t bytes); /* synchronous */
int ggb_host_alloc(size_t bytes, void **hptr); /* pinned host memory */
int ggb_host_free(void *hptr);

/* replaces the profiling probes (src/profiling.h:10-57: _start_recording/_stop_recording/
 * _print_profiling_result) with cudaEvent pairs keyed by the same kind of name strings. */
int ggb_prof_enable(int on);
int ggb_prof_start(const char *name);
int ggb_prof_stop(const char *name);
int ggb_prof_reset(void);
/* total ms and call count accumulated under `name` (resolves pending events; syncs) */
int ggb_prof_get(const char *name, double *total_ms, int64_t *calls);
int ggb_prof_report(char *buf, size_t buflen);

/* CUDA-graph capture of a sequence of ggb_op_* calls (one decode step) */
typedef struct ggb_graph ggb_graph;
int ggb_capture_begin(void);
int ggb_capture_end(ggb_graph **out);
int ggb_graph_launch(ggb_graph *g);
int ggb_graph_free(ggb_graph *g);
int ggb_graph_kernel_count(const ggb_graph *g);

/* ------------------------------------------------------------------ weights
 * A quantised matrix [rows][cols] uploaded ONCE and repacked to the device layout
 * (SoA: one plane of 16 B (Q4_0) / 32 B (Q8_0) quant payloads, one plane of fp16 scales). */
typedef struct ggb_weight ggb_weight;
int ggb_weight_upload(int type, int64_t rows, int64_t cols, const void *host_blocks, int64_t nb01,
                      ggb_weight **out);
/* synthetic matrix generated and quantised on the device (SURVEY.md §8d generator):
 * value(row, col) = synth(seed, tensor_id, row*cols + col) * sigma; bit-identical to the oracle's. */
int ggb_weight_synth(int type, int64_t rows, int64_t cols, uint64_t seed, uint64_t tensor_id, float sigma,
                     ggb_weight **out);
/* rows [row0, row0+nrows) of an existing weight as a new handle (no copy) — TP row shards */
int ggb_weight_slice(const ggb_weight *w, int64_t row0, int64_t nrows, ggb_weight **out);
int ggb_weight_download(const ggb_weight *w, void *host_blocks); /* back to ggml AoS blocks */
int ggb_weight_free(ggb_weight *w);
int64_t ggb_weight_rows(const ggb_weight *w);
int64_t ggb_weight_cols(const ggb_weight *w);
int ggb_weight_type(const ggb_weight *w);
size_t ggb_weight_algorithmic_bytes(const ggb_weight *w); /* rows * cols/32 * (18|34) */
/* f32 vector filled with bias + synth()*sigma (norm weights) */
int ggb_vector_synth(float *dst_dev, int64_t n, uint64_t seed, uint64_t tensor_id, float sigma, float bias);

/* ------------------------------------------------------------------ drop-in mul_mat
 * replaces: extern "C" void mul_mat(...) (src/hpc.h:22-32, src/hpc.cpp:216-277) — same
 * fields, HOST pointers, synchronous: on return all of dst is written.  src0_data/dst_data
 * are src0->data / dst->data of the ggml tensors; vec_dot is gone (the kernel is chosen by
 * src0_type).  wdata = src1 already converted by ggml's INIT stage to rows of `row_size` bytes in the
 * type's vec_dot_type: Q8_0 for Q4_0 / Q8_0 weights, Q8_K for Q4_K / Q6_K (the types the reference's
 * own kernels cover, src/hpc.cpp:132-143), F16 for F16.  Any other src0_type returns
 * GGB_ERR_UNSUPPORTED (the host mirror host/hpc.cpp then runs the caller's vec_dot, as the reference's
 * CPU share does).  src0 is uploaded (Q4_0 / Q8_0: repacked) on first use and cached by
 * (pointer, shape); call ggb_mul_mat_forget(ptr) if the host copy changes. */
int ggb_mul_mat(int64_t ne01, int64_t ne11, int64_t ne12, int64_t nb01, int64_t ne1, int64_t nb1, int64_t nb2,
                size_t row_size, int64_t shared_edge, const void *src0_data, void *dst_data, int src0_type,
                const char *wdata);
int ggb_mul_mat_forget(const void *src0_data);

/* ------------------------------------------------------------------ device operators
 * One entry per graph node kind the builder emits (ref: src/gemma_model.cpp:438-530, 665-747),
 * plus the fused forms the decode step uses.  All pointers are device pointers. */

/* planar Q8_0 activations: per row `cols` int8 followed by cols/32 fp16 scales */
size_t ggb_q8_row_bytes(int64_t cols);
/* K1 — quantize_row_q8_0 (ggml mul_mat INIT stage): T rows of `cols` floats -> planar Q8_0 */
int ggb_op_quantize_q8_0(const float *x, int64_t cols, int64_t T, void *q8);
int ggb_op_q8_to_blocks(const void *q8, int64_t cols, int64_t T, void *blocks);   /* planar -> ggml AoS (34 B) */
int ggb_op_q8_from_blocks(const void *blocks, int64_t cols, int64_t T, int64_t row_size, void *q8);

/* K-quants.  Blocks stay in ggml's AoS layout on the device.
 * ggb_type_row_bytes: bytes of one row of `cols` elements in `type` (0 for an unknown type).
 * ggb_op_quantize_q8_K: [UPSTREAM] quantize_row_q8_K (block: src/kernals.cl:29-34), rows x cols floats ->
 *   rows x cols/256 block_q8_K (292 B), bit-exact.
 * ggb_op_mul_mat_blocks: dst[t][row] = vec_dot(w row, act row t) for type Q4_K / Q6_K (act = Q8_K rows) or F16
 *   (act = fp16 rows); replaces matrix_multiply_q4_K / _q6_K / _f16 (src/kernals.cl:166-265) - which only
 *   compared against the CPU result - with kernels that produce it, in the scalar code's summation order. */
size_t ggb_type_row_bytes(int type, int64_t cols);
int ggb_op_quantize_q8_K(const float *x, int64_t cols, int64_t rows, void *out);
int ggb_op_mul_mat_blocks(int type, const void *w, int64_t rows, int64_t nb01, int64_t K, const void *act,
                          int64_t row_size, int64_t T, float *dst, int64_t dst_stride);
/* Q4_K weights that stay on the device: ggb_op_q4k_repack re-lays a matrix out once (same 144 B per super-block, the
 * nibbles that share a lane-sum index adjacent), ggb_op_mul_mat_q4k_packed is the mat-vec on that layout (same bits as
 * ggb_op_mul_mat_blocks / src/kernals.cl:48-111, ~4x its throughput).  ggb_mul_mat uses it for its cached src0. */
int ggb_op_q4k_repack(const void *raw, int64_t rows, int64_t nb01, int64_t K, void *packed);
int ggb_op_mul_mat_q4k_packed(const void *packed, int64_t rows, int64_t K, const void *act, int64_t row_size, int64_t T,
                              float *dst, int64_t dst_stride);
/* the same for Q6_K: codes expanded to int8 (q - 32) in the transposed arrangement; ggb_q6k_packed_bytes = size of the
 * packed matrix (274 B per super-block, rows padded to 16 B) */
/* the packed mat-vecs on an fp32 activation row x[t][K]: the kernel quantises the row to Q8_K itself while it stages it
 * (ggml's INIT stage of the mul_mat node fused into the COMPUTE stage; same bits as quantize + packed mat-vec) */
int ggb_op_mul_mat_kq_packed_f32(int type, const void *packed, int64_t rows, int64_t K, const float *x, int64_t T, float *dst,
                                 int64_t dst_stride);
/* get_rows on ggml's AoS K-quant blocks (embedding lookup): y[t] = dequantize_row(raw row ids[t]) * scale */
int ggb_op_get_rows_blocks(int type, const void *raw, int64_t n_rows, int64_t nb01, int64_t K, const int32_t *ids, int64_t T,
                           float scale, float *y);
size_t ggb_q6k_packed_bytes(int64_t rows, int64_t K);
int ggb_op_q6k_repack(const void *raw, int64_t rows, int64_t nb01, int64_t K, void *packed);
int ggb_op_mul_mat_q6k_packed(const void *packed, int64_t rows, int64_t K, const void *act, int64_t row_size, int64_t T,
                              float *dst, int64_t dst_stride);

/* K2/K3 — dst[t][row] = dot(w[row], q8[t]) (+ residual[t][row]); dst row pitch dst_stride floats */
int ggb_op_mul_mat_q8(const ggb_weight *w, const void *q8, int64_t T, float *dst, int64_t dst_stride,
                      const float *residual);
/* up to 3 weights sharing one activation row, one launch (Q/K/V) */
int ggb_op_mul_mat_q8_multi(int n, const ggb_weight *const *w, const void *q8, float *const *dst);
/* FFN front half: q8_out = quantize( gelu(gate·x) * (up·x) ); act_f32 optional (dump) */
int ggb_op_ffn_gate_up(const ggb_weight *gate, const ggb_weight *up, const void *q8, void *q8_out,
                       float *act_f32);

/* row-shard form (tensor parallel): this rank's rows write their int8 codes to out_qs and their fp16 block
 * scales to out_d — the caller passes pointers already offset into the gathered planar row */
int ggb_op_ffn_gate_up_shard(const ggb_weight *gate, const ggb_weight *up, const void *q8, void *out_qs, void *out_d,
                             float *act_f32);

/* K5 — rms_norm (+ weight mul when w != NULL) over T rows of E floats; y and/or q8 may be NULL */
int ggb_op_rms_norm(const float *x, const float *w, float eps, int64_t E, int64_t T, float *y, void *q8);
/* element-wise nodes */
int ggb_op_add(const float *a, const float *b, int64_t n, float *y);
int ggb_op_mul(const float *a, const float *b, int64_t n, float *y);
int ggb_op_scale(const float *x, float s, int64_t n, float *y);
int ggb_op_gelu(const float *x, int64_t n, float *y);
/* soft_max_ext(x*scale + mask), max_bias 0 (standalone node, REF semantics except the
 * double sum is a 256-way strided reduction); rows of nc floats; mask may be NULL */
int ggb_op_soft_max_ext(const float *x, const float *mask, float scale, int64_t nc, int64_t nrows, float *y);
/* K6 — rope NEOX on [n_heads][n_dims] rows for T tokens with positions pos[T] (device int32) */
int ggb_op_rope_neox(const float *x, const int32_t *pos, int n_dims, int n_heads, int64_t T, float *y);
/* K8 — get_rows(token_embd, ids) * scale */
int ggb_op_get_rows(const ggb_weight *table, const int32_t *ids, int64_t T, float scale, float *y);
/* K9 — first maximum (strict >); result index written to *out (device int32) */
int ggb_op_argmax(const float *x, int64_t n, int32_t *out);
int ggb_op_cpy_f32_f16(const float *x, int64_t n, void *y_f16);

/* rope tables (host libm recurrence, uploaded once); must be called before rope / decode ops */
int ggb_rope_table_init(int n_ctx, int n_dims, float freq_base);

/* K6 fused: q <- rope(q)*q_scale in place; kcache[pos] <- f16(rope(k)); vcache[pos] <- f16(v).
 * caches are position-major fp16 [n_ctx][Hkv*d]; *pos is a device int32 */
int ggb_op_rope_kv_store(float *q, const float *k, const float *v, const int32_t *pos, int H, int Hkv, int d,
                         float q_scale, void *kcache, void *vcache);
/* the same for T tokens (prefill): q [T][H*d], k / v [T][Hkv*d], pos[T] (device int32); one launch instead of
 * rope(q), scale(q), rope(k) and the two F32->F16 copies */
int ggb_op_rope_kv_store_batch(float *q, const float *k, const float *v, const int32_t *pos, int64_t T, int H, int Hkv,
                               int d, float q_scale, void *kcache, void *vcache);
/* K4 — split-K decode attention over positions [0, *pos]; scratch from ggb_attn_scratch_bytes.
 * out f32 [H*d] and/or q8 (planar, cols = H*d) */
size_t ggb_attn_scratch_bytes(int H, int d, int n_ctx);
int ggb_op_attn_decode(const float *q, const void *kcache, const void *vcache, const int32_t *pos, int H, int Hkv,
                       int d, int n_ctx, void *scratch, float *out, void *q8_out);

/* decode-step glue kept on the device so a captured step can be replayed:
 * state[0] = current token id, state[1] = its position, state[2] = position increment per step
 * (1 for generation; 0 pins n_kv for fixed-context bandwidth measurements) */
int ggb_op_embed_token(const ggb_weight *table, const int32_t *state, float scale, float *y);
/* state[0] <- argmax(logits), state[1] += state[2], history[state[1]] <- token (history may be NULL) */
int ggb_op_sample_greedy(const float *logits, int64_t n, int32_t *state, int32_t *history, void *scratch);
size_t ggb_sample_scratch_bytes(int64_t n);
/* tensor-parallel greedy sampling: each rank publishes (fp32 value, int32 global index) of the first maximum of
 * its logits shard (8 bytes), the candidates are allgathered, then the first maximum over ranks is taken */
int ggb_op_argmax_shard(const float *x, int64_t n, int32_t index_offset, void *cand_out, void *scratch);
int ggb_op_sample_from_candidates(const void *cands, int n_cands, int32_t *state, int32_t *history);

/* ------------------------------------------------------------------ prefill (mul_mat with col_num = T > 1)
 * K10 — C[T][N] = A[T][K] (bf16) x dequant(W[N][K])^T (+ residual): dequant -> bf16 tcgen05 GEMM with TMEM
 * accumulators (csrc/prefill.cu).  replaces: mul_mat with ne11 = T (src/hpc.cpp:216-277) for the graph nodes
 * at src/gemma_model.cpp:692-696, 493, 446-450, 740.  K %% 64 == 0, N %% 32 == 0; workspace (optional,
 * ggb_gemm_workspace_bytes) enables the K-split used for small N. */
size_t ggb_gemm_workspace_bytes(int64_t T, int64_t N, int64_t K);
int ggb_op_gemm_bf16(const ggb_weight *w, const void *a_bf16, int64_t T, float *c, int64_t ldc, const float *resid,
                     void *workspace, size_t workspace_bytes);
/* the same for up to 3 matrices that share the A operand (Q | K | V of src/gemma_model.cpp:692-696, gate | up of
 * :446-447) in ONE launch: c[i] / ldc[i] are the outputs of w[i].  workspace: ggb_gemm_workspace_bytes(T, sum of the
 * row counts rounded up to 256 each, K). */
int ggb_op_gemm_bf16_multi(int n, const ggb_weight *const *w, const void *a_bf16, int64_t T, float *const *c,
                           const int64_t *ldc, void *workspace, size_t workspace_bytes);
int ggb_op_cvt_f32_bf16(const float *x, int64_t n, void *y_bf16);
/* fused producers of the GEMM's bf16 A operand: rms_norm(x) * w, and gelu(gate) * up (graph_build_norm /
 * graph_build_ffn, src/gemma_model.cpp:438-452) */
int ggb_op_rms_norm_bf16(const float *x, const float *w, float eps, int64_t E, int64_t T, void *y_bf16);
int ggb_op_geglu_bf16(const float *gate, const float *up, int64_t n, void *y_bf16);
/* causal attention of T query tokens at positions pos0..pos0+T-1 over the fp16 KV cache (q already roped and
 * scaled, fp32 [T][H*d]); reference node order kq -> soft_max_ext -> kqv (src/gemma_model.cpp:454-489) */
/* out (fp32) and / or out_bf16 (the o-projection's A operand, same rounding as ggb_op_cvt_f32_bf16) may be NULL */
int ggb_op_attn_prefill(const float *q, const void *kcache, const void *vcache, int64_t T, int pos0, int H, int Hkv, int d,
                        float *out, void *out_bf16);

/* ------------------------------------------------------------------ fused decode step
 * The whole per-token node list (ref: src/gemma_model.cpp:665-747) as ONE persistent cooperative
 * kernel: weights and KV stream through a shared-memory ring fed by TMA bulk copies (csrc/decoder.cu).
 * Bit-identical to the ggb_op_* sequence and to the scalar ggml order of the reference's CPU path. */
typedef struct {
    int32_t n_embd, n_layer, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx;
    float eps;
    int32_t pad_;
} ggb_decoder_dims;
typedef struct {
    const float *attn_norm, *ffn_norm;
    const ggb_weight *attn_q, *attn_k, *attn_v, *attn_output, *ffn_gate, *ffn_up, *ffn_down;
    void *k_cache, *v_cache; /* fp16, position-major [n_ctx][n_head_kv*head_dim] */
} ggb_decoder_layer;
typedef struct {
    int32_t *state;   /* [token, pos, pos_step] */
    int32_t *history; /* [n_ctx + 2] or NULL */
    float *sa_out, *l_out; /* [n_layer][n_embd] */
    float *kqv;            /* [n_layer][n_head*head_dim] */
    float *ffn_act;        /* [n_layer][n_ff] or NULL (dump of gelu(gate)*up) */
    float *result_norm;    /* [n_embd] */
    float *logits;         /* [n_vocab] */
} ggb_decoder_io;
typedef struct ggb_decoder ggb_decoder;
/* GGB_ERR_UNSUPPORTED when the shape is outside the fused kernel's envelope (callers then use ggb_op_*) */
int ggb_decoder_create(const ggb_decoder_dims *dims, const ggb_decoder_layer *layers, const ggb_weight *token_embd,
                       const float *output_norm, const ggb_decoder_io *io, ggb_decoder **out);
/* tensor-parallel form (one process per GPU, ggb_comm_init done): the phases that carry the bytes (GATE_UP, DOWN,
 * LOGITS) are split by output rows over `size` ranks and push their flagged 8-byte output words into EVERY rank's
 * buffers with peer stores over NVLink (cudaIpc mappings exchanged at create time) — the all-gather is fused into
 * the producing epilogue, no NCCL call on the token path.  The attention block (QKV, scores, soft-max * V, O) is
 * sharded by KV head when n_head and n_head_kv divide by `size` (each rank's cache then holds only its heads:
 * [n_ctx][ggb_decoder_kv_heads_local * head_dim]) and replicated otherwise.  Bit-identical to the single-GPU kernel. */
int ggb_decoder_kv_heads_local(int n_head, int n_head_kv, int size);
int ggb_decoder_create_tp(const ggb_decoder_dims *dims, const ggb_decoder_layer *layers, const ggb_weight *token_embd,
                          const float *output_norm, const ggb_decoder_io *io, int rank, int size, ggb_decoder **out);
int ggb_decoder_step(ggb_decoder *d); /* enqueue one token; capturable into a ggb_graph */
int ggb_decoder_free(ggb_decoder *d);
/* run ONE traced step: stamps[n_phases][grid][4] %globaltimer values in ns (phase start, after the
 * prologue = inputs arrived, end, end), kinds[n_phases] phase kinds (0 QKV, 1 SCORE, 2 PV, 3 O,
 * 4 GATE_UP, 5 DOWN, 6 LOGITS, 7 SAMPLE) */
int ggb_decoder_trace(ggb_decoder *d, long long *stamps, int32_t *kinds, int max_phases, int *n_phases, int *grid);
/* 1 if the kernel computes the soft-max exp (double exp -> fp32 -> fp16) instead of reading the reference's fp16 table:
 * (2: fp32 expf, 1: double exp; 0: table lookup); enabled only when an exhaustive check at create time found the computed
 * form identical to the table for every non-NaN fp16 input */
int ggb_decoder_exp_computed(const ggb_decoder *d);
/* what THIS rank streams per token: bytes of its pass streams (its row shards plus the phases it runs replicated) and
 * the KV heads whose cache it reads - the per-GPU roofline of a tensor-parallel run is computed from these */
long long ggb_decoder_stream_bytes(const ggb_decoder *d);
int ggb_decoder_kv_heads(const ggb_decoder *d);

/* ------------------------------------------------------------------ multi-GPU (row-sharded TP)
 * one process per GPU; the 128-byte NCCL unique id is created on rank 0 and shipped by the
 * caller (torch.distributed / any transport). */
int ggb_comm_unique_id(void *id128);
int ggb_comm_init(const void *id128, int rank, int world);
int ggb_comm_destroy(void);
int ggb_comm_rank(void);
int ggb_comm_size(void);
/* allgather of `bytes` per rank, in place layout: recv[rank*bytes .. ) */
int ggb_op_allgather(const void *send, void *recv, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* GGML_B200_H */
