// gemma_model.h — host-side model: tensor binding, KV cache, graph builder, greedy sampler.
// Mirrors the reference's class of the same name (ref: src/gemma_model.h:100-236,
// src/gemma_model.cpp) with the graph builder rewired to the C-ABI CUDA layer: instead of ggml
// nodes evaluated by ggml_graph_compute_with_ctx (ref: src/gemma_model.cpp:237) the builder
// enqueues ggb_op_* calls that are captured once into a CUDA graph and replayed per token.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "ggml_b200.h"

typedef int32_t token_id; // ref: src/type.h

enum class inference_stage { PREFILL, DECODE }; // ref: src/gemma_model.h:25-28

// ref: gemma_hyper_param (src/gemma_model.h:30-40), read from GGUF keys at :403-415.  head_dim is
// explicit (the reference derives n_embd/n_head, wrong for Gemma-7B).
struct gemma_config {
    int32_t n_embd, n_layer, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx;
    float eps, rope_base;
    int32_t wtype;
    int32_t pad_;
    uint64_t seed;
    float sigma_embd, sigma_mat;
};

// tensor ids in the reference's binding order (ref: src/gemma_model.cpp:159-194)
enum { GM_T_TOKEN_EMBD = 0, GM_T_OUTPUT_NORM = 1, GM_T_LAYER0 = 2, GM_T_PER_LAYER = 9 };
enum { GM_L_ATTN_NORM = 0, GM_L_ATTN_Q, GM_L_ATTN_K, GM_L_ATTN_V, GM_L_ATTN_OUT,
       GM_L_FFN_NORM, GM_L_FFN_GATE, GM_L_FFN_UP, GM_L_FFN_DOWN };

struct gemma_layer { // ref: gemma_layer (src/gemma_model.h:42-50 region)
    float *attn_norm = nullptr, *ffn_norm = nullptr;
    ggb_weight *attn_q = nullptr, *attn_k = nullptr, *attn_v = nullptr, *attn_output = nullptr;
    ggb_weight *ffn_gate = nullptr, *ffn_up = nullptr, *ffn_down = nullptr;
    // tensor-parallel row shards of the above (views, no copy); null when tp_size == 1
    ggb_weight *sh_q = nullptr, *sh_k = nullptr, *sh_v = nullptr, *sh_o = nullptr;
    ggb_weight *sh_gate = nullptr, *sh_up = nullptr, *sh_down = nullptr;
};

// Row-sharded tensor parallelism: every weight mat-vec is split by OUTPUT ROWS — the axis the
// reference already splits between its CPU threads and the GPU (ref: src/hpc.cpp:245-269) — so each output
// element is computed by exactly one rank in the same K-order and sharded results are bit-identical to
// 1-GPU results.  Pure host logic, testable without a GPU.
struct tp_range {
    int64_t row0, nrows;
};
struct tp_plan {
    int32_t rank, size, valid;
    int32_t head0, heads;        // query heads owned
    int32_t kv_head0, kv_heads;  // KV heads held in this rank's cache
    int32_t kv_replicated;       // MQA (1 KV head): every rank keeps the whole (tiny) K/V projection and cache
    tp_range q, k, v, o, gate_up, down, logits;
};
tp_plan make_tp_plan(const gemma_config &cfg, int rank, int size);

// ref: gemma_tokenizer (src/gemma_model.h; print_token / print_tokens src/gemma_model.cpp:749-793), filled from the
// GGUF vocabulary by host/gguf_loader.cpp
struct gemma_tokenizer;

class gemma_model {
  public:
    explicit gemma_model(const gemma_config &cfg);
    ~gemma_model();
    // ref: load_model_from_file (src/gemma_model.cpp:19-51): hyper-parameters from the GGUF header (:403-415), the
    // 2 + 9 L tensors bound by the reference's names (:159-194) and uploaded, the tokenizer vocabulary.  n_ctx = 0
    // takes gemma.context_length.  Returns nullptr and sets *err on failure.
    static gemma_model *load_model_from_file(const char *file_path, int32_t n_ctx, std::string *err);
    const gemma_tokenizer *tokenizer() const { return tokenizer_; }
    // ref: begin_one_round_inference (:548-575): prefill, then decode until the EOS token or max_tokens / the context
    // is reached.  `text` (optional) receives what the reference prints: the detokenised prompt and every new token.
    int begin_one_round_inference(const std::vector<token_id> &prompt, int max_tokens, std::vector<token_id> &out,
                                  std::string *text, double *prefill_ms, double *decode_ms);

    // call before the weights exist; NCCL (ggb_comm_init) must already be up for size > 1
    int init_tensor_parallel(int rank, int size);
    const tp_plan &plan() const { return plan_; }
    // weights: synthetic (generated on the device) or uploaded tensor by tensor
    int init_synthetic_weights();
    int set_tensor(int tensor_id, const void *host_data, int64_t nbytes);
    // K-quant files (the author's q4_k_m GGUF: matrices Q4_K / Q6_K, norms F32): a matrix of type GGB_TYPE_Q4_K or
    // GGB_TYPE_Q6_K puts the model into K-QUANT MODE - every matrix must then be a K-quant; the token step is lowered
    // operator by operator (quantize_row_q8_K, the packed K-quant mat-vecs, get_rows on the blocks: csrc/kquants.cu) and
    // captured into a CUDA graph like the Q4_0 / Q8_0 per-operator form.  Other types behave like set_tensor.
    int set_tensor_typed(int tensor_id, int type, const void *host_data, int64_t nbytes);
    bool kquant_mode() const { return kq_mode_; }
    int init_kv_cache();                       // ref: :361-401
    int fill_kv_cache_synthetic(uint64_t seed); // bandwidth sweeps
    // ref: :665-747.  fused = true lowers the token step to the persistent decode kernel
    // (ggb_decoder_*); fused = false emits one ggb_op_* per (fused group of) graph node(s).
    int build_compute_graph(bool capture_acts, bool fused = true);
    bool is_fused() const { return decoder_ != nullptr; }
    ggb_decoder *decoder() const { return decoder_; }

    // ref: inference() (:231-286): appends the sampled token to `input`
    int inference(std::vector<token_id> &input, inference_stage stage);
    // ref: begin_one_round_inference() (:548-575)
    int generate(const std::vector<token_id> &prompt, int n_decode, std::vector<token_id> &out, double *prefill_ms,
                 double *decode_ms);

    // Batched prefill of T tokens at positions pos0 .. pos0+T-1 through the dequant -> bf16 tcgen05 GEMMs
    // (ggb_op_gemm_bf16): the reference's PREFILL stage evaluates the same graph with col_num = T
    // (ref: inference(), src/gemma_model.cpp:231-286; logits for all T rows, :740).  Fills the KV cache, leaves
    // the decode state at (next token, pos0 + T) and returns the greedy token of the last position.
    // Tolerance-exact (bf16 activations), unlike the token-by-token PREFILL above which is bit-exact.
    int prefill(const token_id *tokens, int T, int32_t pos0, token_id *next);
    int get_prefill_logits(int row, float *host) const; // row in [0, T) of the last prefill
    double last_prefill_gemm_flops() const { return pf_flops_; }

    // one decode step through host buffers: H2D (token,pos), graph, D2H next token
    int decode_step(token_id token, int32_t pos, token_id *next);
    // device-resident loop: `steps` graph launches, no host round trip; pos_step 0 pins n_kv
    int set_state(token_id token, int32_t pos, int32_t pos_step);
    int run_steps(int steps);
    int run_steps_profiled(int steps); // eager, cudaEvent probes per op class
    int get_state(token_id *token, int32_t *pos);

    int get_logits(float *host) const;
    int get_capture(const std::string &name, int layer, float *host, int64_t n) const;
    int get_kv(int layer, void *k_host, void *v_host) const;
    int kernels_per_step() const { return graph_ ? ggb_graph_kernel_count(graph_) : (decoder_ ? 1 : 0); }
    size_t weight_bytes() const { return weight_bytes_; }
    const gemma_config &config() const { return cfg_; }

  private:
    int alloc_buffers();
    int emit_decode_step(); // the op sequence of one token
    int emit_decode_step_tp(); // same, row-sharded over tp_size ranks with allgathers at the layer boundaries
    int emit_decode_step_kq(); // same node list on K-quant weights (Q8_K activations)
    int make_shards();
    int launch_step();      // graph replay, or a direct cooperative launch
    void free_all();

    gemma_config cfg_;
    gemma_tokenizer *tokenizer_ = nullptr;
    ggb_weight *token_embd_ = nullptr; // also the output projection (tied, ref: :163)
    float *output_norm_ = nullptr;
    std::vector<gemma_layer> layers_;
    std::vector<void *> k_cache_, v_cache_; // ref: kv_cache (src/gemma_model.h:83-97)
    size_t weight_bytes_ = 0;

    // K-quant mode: matrices as packed K-quant copies (+ the raw token_embd blocks for get_rows)
    struct kq_mat {
        int type = 0;
        void *raw = nullptr, *packed = nullptr;
        int64_t rows = 0, cols = 0;
    };
    bool kq_mode_ = false;
    kq_mat kq_embd_;
    std::vector<kq_mat> kq_layers_; // [n_layer][7]: q, k, v, o, gate, up, down
    float *kq_xn_ = nullptr, *kq_gate_ = nullptr, *kq_up_ = nullptr;
    void *kq_q8_ = nullptr; // one Q8_K activation row (longest K)
    int kq_mv(const kq_mat &w, const float *x, float *dst); // the packed mat-vec on the fp32 row x (quantised to Q8_K in its staging)

    // device-resident step state: [token, pos, pos_step]
    int32_t *state_ = nullptr, *history_ = nullptr;
    int32_t *h_state_ = nullptr; // pinned
    int64_t pos_hint_ = 0;       // host mirror of the device position / step (run_steps refuses to leave the context)
    int32_t pos_step_hint_ = 0;
    // activations
    float *x0_ = nullptr, *q_ = nullptr, *k_ = nullptr, *v_ = nullptr, *logits_ = nullptr, *result_norm_ = nullptr;
    float *sa_out_ = nullptr, *l_out_ = nullptr, *kqv_ = nullptr, *ffn_act_ = nullptr; // per layer slots
    void *act_e_ = nullptr, *act_q_ = nullptr, *act_f_ = nullptr;                      // planar Q8_0
    void *attn_scratch_ = nullptr, *sample_scratch_ = nullptr;
    bool capture_acts_ = false;
    ggb_graph *graph_ = nullptr;
    ggb_decoder *decoder_ = nullptr;
    bool built_ = false;
    tp_plan plan_{};
    // multi-GPU lowering: true = the persistent kernel with flagged peer stores over NVLink (ggb_decoder_create_tp;
    // KV cache in the full [n_ctx][Hkv*d] layout, each rank fills the chunks it owns); false = per-operator
    // kernels + NCCL allgathers (KV cache sharded by head).  GGB_TP_FUSED=0 selects the latter.
    bool tp_fused_ = true;
    // KV heads this rank's cache holds: its own heads when attention is sharded by head, all of them when the attention
    // block is replicated (MQA under the fused lowering)
    int kv_heads_alloc() const {
        return (plan_.size > 1 && tp_fused_) ? ggb_decoder_kv_heads_local(cfg_.n_head, cfg_.n_head_kv, plan_.size) : plan_.kv_heads;
    }
    ggb_weight *sh_logits_ = nullptr;
    float *kqv_local_ = nullptr;
    void *cand_ = nullptr; // [tp_size] (value, index) candidates of the sharded arg-max
    // prefill buffers, sized for pf_T_ tokens
    int alloc_prefill(int T);
    void free_prefill();
    int pf_T_ = 0, pf_last_T_ = 0;
    double pf_flops_ = 0.0;
    int32_t *pf_ids_ = nullptr, *pf_pos_ = nullptr;
    float *pf_x_ = nullptr, *pf_xn_ = nullptr, *pf_q_ = nullptr, *pf_k_ = nullptr, *pf_v_ = nullptr, *pf_attn_ = nullptr;
    float *pf_sa_ = nullptr, *pf_g_ = nullptr, *pf_u_ = nullptr, *pf_logits_ = nullptr;
    void *pf_a_ = nullptr, *pf_ws_ = nullptr; // bf16 A operand, K-split workspace
    size_t pf_ws_bytes_ = 0;
    int64_t pf_ws_n_ = 0; // widest N the K-split workspace covers
};
