// gemma_model.cpp — host side of the decode path.  Mirrors ref: src/gemma_model.cpp
// (tensor binding :159-194, KV cache :361-401 / :428-436, graph builder :438-530 / :665-747,
// inference :231-286, greedy round :548-575) with every node lowered to the C-ABI CUDA layer.
#include "gemma_model.h"
#include "gguf_loader.h"
#include <cstdlib>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>

#define GM_TRY(call)          \
    do {                      \
        int rc__ = (call);    \
        if (rc__) return rc__; \
    } while (0)

static void kq_free(void *&p) {
    if (p) ggb_buffer_free(p);
    p = nullptr;
}

static constexpr float kNormSigma = 0.02f; // norm weights: 1 + N(0, 0.02^2)

gemma_model::gemma_model(const gemma_config &cfg) : cfg_(cfg) {
    layers_.resize((size_t)cfg.n_layer);
    plan_ = make_tp_plan(cfg, 0, 1);
}

// ---------------------------------------------------------------- tensor-parallel plan (host logic only)
tp_plan make_tp_plan(const gemma_config &c, int rank, int size) {
    tp_plan p{};
    p.rank = rank;
    p.size = size;
    const int64_t d = c.head_dim, E = c.n_embd, F = c.n_ff, V = c.n_vocab;
    if (size < 1 || rank < 0 || rank >= size) return p;
    if (c.n_head % size || E % size || V % size || F % (32 * size)) return p; // GeGLU+quantise works on 32-row blocks
    p.kv_replicated = (c.n_head_kv % size) != 0;
    if (p.kv_replicated && c.n_head_kv != 1) return p; // only MQA is replicated
    p.heads = c.n_head / size;
    p.head0 = rank * p.heads;
    p.kv_heads = p.kv_replicated ? c.n_head_kv : c.n_head_kv / size;
    p.kv_head0 = p.kv_replicated ? 0 : rank * p.kv_heads;
    p.q = {p.head0 * d, p.heads * d};
    p.k = {p.kv_head0 * d, p.kv_heads * d};
    p.v = p.k;
    p.o = {rank * (E / size), E / size};
    p.gate_up = {rank * (F / size), F / size};
    p.down = p.o;
    p.logits = {rank * (V / size), V / size};
    p.valid = 1;
    return p;
}

int gemma_model::init_tensor_parallel(int rank, int size) {
    if (token_embd_) return GGB_ERR_ARG; // must precede weight creation
    tp_plan p = make_tp_plan(cfg_, rank, size);
    if (const char *e = getenv("GGB_TP_FUSED")) tp_fused_ = atoi(e) != 0;
    if (!p.valid && tp_fused_ && size >= 1 && rank >= 0 && rank < size && cfg_.n_vocab % (2 * size) == 0) {
        // the persistent kernel splits every phase by row pairs / 32-row groups, whatever the head count: only
        // the logits shard boundaries (V / size rows per rank, used by get_logits) come from the plan
        p.kv_heads = cfg_.n_head_kv;
        p.logits = {(int64_t)rank * (cfg_.n_vocab / size), cfg_.n_vocab / size};
        p.valid = 1;
    }
    if (!p.valid) return GGB_ERR_ARG;
    if (size > 1 && (ggb_comm_size() != size || ggb_comm_rank() != rank)) return GGB_ERR_ARG;
    plan_ = p;
    return 0;
}

int gemma_model::make_shards() {
    if (plan_.size == 1) return 0;
    auto sl = [&](const ggb_weight *w, const tp_range &r, ggb_weight **out) -> int {
        ggb_weight_free(*out);
        *out = nullptr;
        // a plan made for the persistent kernel only (head count not divisible by the world size) has no
        // per-operator ranges: that lowering is not built, the kernel slices its own pass streams
        if (r.nrows == 0) return 0;
        return ggb_weight_slice(w, r.row0, r.nrows, out);
    };
    for (auto &l : layers_) {
        GM_TRY(sl(l.attn_q, plan_.q, &l.sh_q));
        GM_TRY(sl(l.attn_k, plan_.k, &l.sh_k));
        GM_TRY(sl(l.attn_v, plan_.v, &l.sh_v));
        GM_TRY(sl(l.attn_output, plan_.o, &l.sh_o));
        GM_TRY(sl(l.ffn_gate, plan_.gate_up, &l.sh_gate));
        GM_TRY(sl(l.ffn_up, plan_.gate_up, &l.sh_up));
        GM_TRY(sl(l.ffn_down, plan_.down, &l.sh_down));
    }
    return sl(token_embd_, plan_.logits, &sh_logits_);
}

gemma_model::~gemma_model() {
    free_all();
    delete tokenizer_;
}

gemma_model *gemma_model::load_model_from_file(const char *file_path, int32_t n_ctx, std::string *err) {
    gguf_file f;
    if (f.open(file_path) != 0) {
        if (err) *err = f.error();
        return nullptr;
    }
    gemma_config cfg;
    if (gguf_gemma_config(f, n_ctx, &cfg, err) != 0) return nullptr;
    gemma_model *m = new gemma_model(cfg);
    if (gguf_bind_tensors(f, *m, err) != 0) {
        delete m;
        return nullptr;
    }
    m->tokenizer_ = new gemma_tokenizer();
    if (gguf_gemma_tokenizer(f, m->tokenizer_) != 0) { // a file without a vocabulary still runs: ids only
        delete m->tokenizer_;
        m->tokenizer_ = nullptr;
    }
    return m;
}

int gemma_model::begin_one_round_inference(const std::vector<token_id> &prompt, int max_tokens, std::vector<token_id> &out,
                                           std::string *text, double *prefill_ms, double *decode_ms) {
    using clk = std::chrono::high_resolution_clock;
    if (prompt.empty() || (int)prompt.size() >= cfg_.n_ctx) return GGB_ERR_ARG;
    const token_id eos = tokenizer_ ? tokenizer_->special_eos_id : -1;
    std::vector<token_id> input = prompt;
    if (text && tokenizer_) *text = tokenizer_->tokens_text(input); // ref: print_tokens(input)
    GM_TRY(init_kv_cache());
    const auto t0 = clk::now();
    GM_TRY(inference(input, inference_stage::PREFILL));
    const auto t1 = clk::now();
    // ref :559-563: while (last != eos && size < DEFAULT_TOKEN_NUM) { print_token(last); inference(DECODE); }
    while (input.back() != eos && (int)(input.size() - prompt.size()) < max_tokens && (int)input.size() < cfg_.n_ctx) {
        if (text && tokenizer_) *text += tokenizer_->token_text(input.back());
        GM_TRY(inference(input, inference_stage::DECODE));
    }
    const auto t2 = clk::now();
    out.assign(input.begin() + (long)prompt.size(), input.end());
    if (prefill_ms) *prefill_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (decode_ms) *decode_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    return 0;
}

void gemma_model::free_all() {
    ggb_sync();
    free_prefill();
    ggb_graph_free(graph_);
    graph_ = nullptr;
    ggb_decoder_free(decoder_);
    decoder_ = nullptr;
    ggb_weight_free(token_embd_);
    token_embd_ = nullptr;
    ggb_weight_free(sh_logits_);
    sh_logits_ = nullptr;
    for (auto &l : layers_) {
        ggb_weight *ws[] = {l.sh_q, l.sh_k, l.sh_v, l.sh_o, l.sh_gate, l.sh_up, l.sh_down,
                            l.attn_q, l.attn_k, l.attn_v, l.attn_output, l.ffn_gate, l.ffn_up, l.ffn_down};
        for (auto *w : ws) ggb_weight_free(w);
        if (l.attn_norm) ggb_buffer_free(l.attn_norm);
        if (l.ffn_norm) ggb_buffer_free(l.ffn_norm);
        l = gemma_layer();
    }
    void *bufs[] = {output_norm_, state_, history_, x0_, q_, k_, v_, logits_, result_norm_, sa_out_, l_out_, kqv_,
                    ffn_act_, act_e_, act_q_, act_f_, attn_scratch_, sample_scratch_, kqv_local_, cand_};
    kqv_local_ = nullptr;
    cand_ = nullptr;
    for (void *b : bufs)
        if (b) ggb_buffer_free(b);
    output_norm_ = nullptr; state_ = history_ = nullptr; x0_ = q_ = k_ = v_ = logits_ = result_norm_ = nullptr;
    sa_out_ = l_out_ = kqv_ = ffn_act_ = nullptr; act_e_ = act_q_ = act_f_ = attn_scratch_ = sample_scratch_ = nullptr;
    kq_free(kq_embd_.raw);
    kq_free(kq_embd_.packed);
    for (kq_mat &w : kq_layers_) {
        kq_free(w.raw);
        kq_free(w.packed);
    }
    kq_layers_.clear();
    kq_embd_ = kq_mat();
    kq_mode_ = false;
    void *kqb[] = {kq_xn_, kq_gate_, kq_up_, kq_q8_};
    for (void *b : kqb)
        if (b) ggb_buffer_free(b);
    kq_xn_ = kq_gate_ = kq_up_ = nullptr;
    kq_q8_ = nullptr;
    for (void *b : k_cache_) ggb_buffer_free(b);
    for (void *b : v_cache_) ggb_buffer_free(b);
    k_cache_.clear();
    v_cache_.clear();
    if (h_state_) ggb_host_free(h_state_);
    h_state_ = nullptr;
}

// ref: composite_model / load_tensors (src/gemma_model.cpp:159-194): token_embd (also the output
// projection), output_norm, and nine tensors per block.  Here every matrix is generated and
// quantised ON the device with the shared counter-based generator (seed, tensor id).
int gemma_model::init_synthetic_weights() {
    if (kq_mode_ || (cfg_.wtype != GGB_TYPE_Q4_0 && cfg_.wtype != GGB_TYPE_Q8_0)) return GGB_ERR_UNSUPPORTED; // the device generator makes Q4_0 / Q8_0
    const gemma_config &c = cfg_;
    const int64_t E = c.n_embd, F = c.n_ff, V = c.n_vocab;
    const int64_t QD = (int64_t)c.n_head * c.head_dim, KD = (int64_t)c.n_head_kv * c.head_dim;
    weight_bytes_ = 0;
    auto mat = [&](int tid, int64_t rows, int64_t cols, ggb_weight **out) -> int {
        const float sigma = tid == GM_T_TOKEN_EMBD ? c.sigma_embd : c.sigma_mat;
        GM_TRY(ggb_weight_synth(c.wtype, rows, cols, c.seed, (uint64_t)tid, sigma, out));
        weight_bytes_ += ggb_weight_algorithmic_bytes(*out);
        return 0;
    };
    auto vec = [&](int tid, int64_t n, float **out) -> int {
        GM_TRY(ggb_buffer_alloc((size_t)n * 4, (void **)out));
        return ggb_vector_synth(*out, n, c.seed, (uint64_t)tid, kNormSigma, 1.0f);
    };
    GM_TRY(mat(GM_T_TOKEN_EMBD, V, E, &token_embd_));
    GM_TRY(vec(GM_T_OUTPUT_NORM, E, &output_norm_));
    for (int il = 0; il < c.n_layer; il++) {
        const int b = GM_T_LAYER0 + GM_T_PER_LAYER * il;
        gemma_layer &l = layers_[(size_t)il];
        GM_TRY(vec(b + GM_L_ATTN_NORM, E, &l.attn_norm));
        GM_TRY(mat(b + GM_L_ATTN_Q, QD, E, &l.attn_q));
        GM_TRY(mat(b + GM_L_ATTN_K, KD, E, &l.attn_k));
        GM_TRY(mat(b + GM_L_ATTN_V, KD, E, &l.attn_v));
        GM_TRY(mat(b + GM_L_ATTN_OUT, E, QD, &l.attn_output));
        GM_TRY(vec(b + GM_L_FFN_NORM, E, &l.ffn_norm));
        GM_TRY(mat(b + GM_L_FFN_GATE, F, E, &l.ffn_gate));
        GM_TRY(mat(b + GM_L_FFN_UP, F, E, &l.ffn_up));
        GM_TRY(mat(b + GM_L_FFN_DOWN, E, F, &l.ffn_down));
    }
    return ggb_sync();
}

// upload one tensor from host memory (ggml AoS blocks for matrices, f32 for norms)
int gemma_model::set_tensor(int tensor_id, const void *host_data, int64_t nbytes) {
    const gemma_config &c = cfg_;
    // the captured graph, the persistent kernel's pass streams and the tensor-parallel shard views all hold pointers
    // into the weights: replacing one invalidates them (build_compute_graph must be called again)
    if (built_ || graph_ || decoder_) {
        ggb_sync();
        ggb_graph_free(graph_);
        graph_ = nullptr;
        ggb_decoder_free(decoder_);
        decoder_ = nullptr;
        built_ = false;
    }
    const int64_t E = c.n_embd, F = c.n_ff, V = c.n_vocab;
    const int64_t QD = (int64_t)c.n_head * c.head_dim, KD = (int64_t)c.n_head_kv * c.head_dim;
    auto mat = [&](ggb_weight **slot, int64_t rows, int64_t cols) -> int {
        if (kq_mode_ || (c.wtype != GGB_TYPE_Q4_0 && c.wtype != GGB_TYPE_Q8_0)) return GGB_ERR_ARG; // K-quant matrices: set_tensor_typed
        const int64_t need = rows * (cols / 32) * (c.wtype == GGB_TYPE_Q4_0 ? 18 : 34);
        if (nbytes != need) return GGB_ERR_ARG;
        if (*slot) weight_bytes_ -= ggb_weight_algorithmic_bytes(*slot);
        ggb_weight_free(*slot);
        *slot = nullptr;
        GM_TRY(ggb_weight_upload(c.wtype, rows, cols, host_data, 0, slot));
        weight_bytes_ += ggb_weight_algorithmic_bytes(*slot);
        return 0;
    };
    auto vec = [&](float **slot, int64_t n) -> int {
        if (nbytes != n * 4) return GGB_ERR_ARG;
        if (!*slot) GM_TRY(ggb_buffer_alloc((size_t)n * 4, (void **)slot));
        GM_TRY(ggb_upload(*slot, host_data, (size_t)n * 4));
        return ggb_sync();
    };
    if (tensor_id == GM_T_TOKEN_EMBD) return mat(&token_embd_, V, E);
    if (tensor_id == GM_T_OUTPUT_NORM) return vec(&output_norm_, E);
    const int rel = tensor_id - GM_T_LAYER0;
    if (rel < 0 || rel >= GM_T_PER_LAYER * c.n_layer) return GGB_ERR_ARG;
    gemma_layer &l = layers_[(size_t)(rel / GM_T_PER_LAYER)];
    switch (rel % GM_T_PER_LAYER) {
    case GM_L_ATTN_NORM: return vec(&l.attn_norm, E);
    case GM_L_ATTN_Q: return mat(&l.attn_q, QD, E);
    case GM_L_ATTN_K: return mat(&l.attn_k, KD, E);
    case GM_L_ATTN_V: return mat(&l.attn_v, KD, E);
    case GM_L_ATTN_OUT: return mat(&l.attn_output, E, QD);
    case GM_L_FFN_NORM: return vec(&l.ffn_norm, E);
    case GM_L_FFN_GATE: return mat(&l.ffn_gate, F, E);
    case GM_L_FFN_UP: return mat(&l.ffn_up, F, E);
    case GM_L_FFN_DOWN: return mat(&l.ffn_down, E, F);
    }
    return GGB_ERR_ARG;
}

int gemma_model::set_tensor_typed(int tensor_id, int type, const void *host_data, int64_t nbytes) {
    if (type != GGB_TYPE_Q4_K && type != GGB_TYPE_Q6_K) {
        if (kq_mode_ && tensor_id != GM_T_OUTPUT_NORM && (tensor_id < GM_T_LAYER0 || ((tensor_id - GM_T_LAYER0) % GM_T_PER_LAYER != GM_L_ATTN_NORM &&
                                                                                       (tensor_id - GM_T_LAYER0) % GM_T_PER_LAYER != GM_L_FFN_NORM)))
            return GGB_ERR_ARG; // a K-quant model takes K-quant matrices only
        return set_tensor(tensor_id, host_data, nbytes);
    }
    const gemma_config &c = cfg_;
    if (plan_.size > 1) return GGB_ERR_UNSUPPORTED; // K-quant mode is single-GPU
    if (token_embd_) return GGB_ERR_ARG;            // a model is either Q4_0 / Q8_0 or K-quant, not both
    for (const gemma_layer &l : layers_)
        if (l.attn_q || l.attn_k || l.attn_v || l.attn_output || l.ffn_gate || l.ffn_up || l.ffn_down) return GGB_ERR_ARG;
    if (built_ || graph_ || decoder_) {
        ggb_sync();
        ggb_graph_free(graph_);
        graph_ = nullptr;
        ggb_decoder_free(decoder_);
        decoder_ = nullptr;
        built_ = false;
    }
    const int64_t E = c.n_embd, F = c.n_ff, V = c.n_vocab;
    const int64_t QD = (int64_t)c.n_head * c.head_dim, KD = (int64_t)c.n_head_kv * c.head_dim;
    if (kq_layers_.empty()) kq_layers_.resize((size_t)c.n_layer * 7);
    kq_mat *slot = nullptr;
    int64_t rows = 0, cols = 0;
    bool keep_raw = false;
    if (tensor_id == GM_T_TOKEN_EMBD) {
        slot = &kq_embd_; rows = V; cols = E; keep_raw = true;
    } else {
        const int rel = tensor_id - GM_T_LAYER0;
        if (rel < 0 || rel >= GM_T_PER_LAYER * c.n_layer) return GGB_ERR_ARG;
        const int il = rel / GM_T_PER_LAYER;
        int k = -1;
        switch (rel % GM_T_PER_LAYER) {
        case GM_L_ATTN_Q: k = 0; rows = QD; cols = E; break;
        case GM_L_ATTN_K: k = 1; rows = KD; cols = E; break;
        case GM_L_ATTN_V: k = 2; rows = KD; cols = E; break;
        case GM_L_ATTN_OUT: k = 3; rows = E; cols = QD; break;
        case GM_L_FFN_GATE: k = 4; rows = F; cols = E; break;
        case GM_L_FFN_UP: k = 5; rows = F; cols = E; break;
        case GM_L_FFN_DOWN: k = 6; rows = E; cols = F; break;
        default: return GGB_ERR_ARG; // norms are F32
        }
        slot = &kq_layers_[(size_t)il * 7 + (size_t)k];
    }
    if (cols % 256 != 0) return GGB_ERR_ARG;
    const int64_t row_bytes = (int64_t)ggb_type_row_bytes(type, cols);
    if (nbytes != rows * row_bytes) return GGB_ERR_ARG;
    if (slot->packed) weight_bytes_ -= (size_t)(slot->rows * (int64_t)ggb_type_row_bytes(slot->type, slot->cols));
    kq_free(slot->raw);
    kq_free(slot->packed);
    GM_TRY(ggb_buffer_alloc((size_t)nbytes, &slot->raw));
    GM_TRY(ggb_upload(slot->raw, host_data, (size_t)nbytes));
    const size_t pbytes = type == GGB_TYPE_Q4_K ? (size_t)nbytes : ggb_q6k_packed_bytes(rows, cols);
    GM_TRY(ggb_buffer_alloc(pbytes, &slot->packed));
    GM_TRY(type == GGB_TYPE_Q4_K ? ggb_op_q4k_repack(slot->raw, rows, row_bytes, cols, slot->packed)
                                  : ggb_op_q6k_repack(slot->raw, rows, row_bytes, cols, slot->packed));
    GM_TRY(ggb_sync());
    if (!keep_raw) kq_free(slot->raw);
    slot->type = type; slot->rows = rows; slot->cols = cols;
    weight_bytes_ += (size_t)nbytes;
    kq_mode_ = true;
    return 0;
}

// ref: init_kv_cache (:361-401): per layer one F16 K and one F16 V tensor, zeroed.  Device layout is
// position-major [n_ctx][n_head_kv*head_dim] for both (see csrc/attention.cu).
int gemma_model::init_kv_cache() {
    const size_t bytes = (size_t)cfg_.n_ctx * kv_heads_alloc() * cfg_.head_dim * 2;
    for (int il = 0; il < cfg_.n_layer; il++) {
        void *k = nullptr, *v = nullptr;
        if ((int)k_cache_.size() <= il) {
            GM_TRY(ggb_buffer_alloc(bytes, &k));
            GM_TRY(ggb_buffer_alloc(bytes, &v));
            k_cache_.push_back(k);
            v_cache_.push_back(v);
        }
        GM_TRY(ggb_buffer_memset(k_cache_[(size_t)il], 0, bytes));
        GM_TRY(ggb_buffer_memset(v_cache_[(size_t)il], 0, bytes));
    }
    GM_TRY(ggb_rope_table_init(cfg_.n_ctx, cfg_.head_dim, cfg_.rope_base));
    return ggb_sync();
}

int gemma_model::fill_kv_cache_synthetic(uint64_t seed) {
    const int64_t n = (int64_t)cfg_.n_ctx * kv_heads_alloc() * cfg_.head_dim;
    float *tmp = nullptr;
    GM_TRY(ggb_buffer_alloc((size_t)n * 4, (void **)&tmp));
    for (int il = 0; il < cfg_.n_layer; il++) {
        GM_TRY(ggb_vector_synth(tmp, n, seed, (uint64_t)(2 * il), 1.0f, 0.0f));
        GM_TRY(ggb_op_cpy_f32_f16(tmp, n, k_cache_[(size_t)il]));
        GM_TRY(ggb_vector_synth(tmp, n, seed, (uint64_t)(2 * il + 1), 1.0f, 0.0f));
        GM_TRY(ggb_op_cpy_f32_f16(tmp, n, v_cache_[(size_t)il]));
    }
    GM_TRY(ggb_sync());
    return ggb_buffer_free(tmp);
}

int gemma_model::alloc_buffers() {
    const gemma_config &c = cfg_;
    const size_t E = (size_t)c.n_embd, F = (size_t)c.n_ff, V = (size_t)c.n_vocab, L = (size_t)c.n_layer;
    const size_t QD = (size_t)c.n_head * c.head_dim, KD = (size_t)c.n_head_kv * c.head_dim;
    if (state_) return 0;
    GM_TRY(ggb_buffer_alloc(16, (void **)&state_));
    GM_TRY(ggb_buffer_alloc(((size_t)c.n_ctx + 2) * 4, (void **)&history_));
    GM_TRY(ggb_host_alloc(32, (void **)&h_state_));
    GM_TRY(ggb_buffer_alloc(E * 4, (void **)&x0_));
    GM_TRY(ggb_buffer_alloc(QD * 4, (void **)&q_));
    GM_TRY(ggb_buffer_alloc(KD * 4, (void **)&k_));
    GM_TRY(ggb_buffer_alloc(KD * 4, (void **)&v_));
    GM_TRY(ggb_buffer_alloc(V * 4, (void **)&logits_));
    GM_TRY(ggb_buffer_alloc(E * 4, (void **)&result_norm_));
    GM_TRY(ggb_buffer_alloc(L * E * 4, (void **)&sa_out_));
    GM_TRY(ggb_buffer_alloc(L * E * 4, (void **)&l_out_));
    GM_TRY(ggb_buffer_alloc(L * QD * 4, (void **)&kqv_));
    GM_TRY(ggb_buffer_alloc(L * F * 4, (void **)&ffn_act_));
    GM_TRY(ggb_buffer_alloc(ggb_q8_row_bytes((int64_t)E), &act_e_));
    GM_TRY(ggb_buffer_alloc(ggb_q8_row_bytes((int64_t)QD), &act_q_));
    GM_TRY(ggb_buffer_alloc(ggb_q8_row_bytes((int64_t)F), &act_f_));
    GM_TRY(ggb_buffer_alloc(ggb_attn_scratch_bytes(c.n_head, c.head_dim, c.n_ctx), &attn_scratch_));
    GM_TRY(ggb_buffer_alloc(ggb_sample_scratch_bytes((int64_t)V), &sample_scratch_));
    GM_TRY(ggb_buffer_alloc(QD * 4, (void **)&kqv_local_));
    GM_TRY(ggb_buffer_alloc(8 * (size_t)(plan_.size > 0 ? plan_.size : 1), &cand_));
    GM_TRY(ggb_buffer_memset(state_, 0, 16));
    GM_TRY(ggb_buffer_memset(history_, 0, ((size_t)c.n_ctx + 2) * 4));
    return ggb_sync();
}

// The op sequence of ONE token — the reference's per-layer node list (SURVEY.md §3.3,
// ref: src/gemma_model.cpp:688-740) with adjacent nodes fused:
//   norm+mul+quantize | Q,K,V mat-vecs | rope+scale+KV store | kq/softmax/kqv split-K (+quantize)
//   | o-proj + residual | norm+mul+quantize | gate,up mat-vecs + gelu*up + quantize | down + residual
// profiling probe with the reference's name keys (ref: notes.md:68-91 "op_mul_mat", "op_rms_norm", ...;
// src/profiling.h:14-18 profile(str, code)); a no-op unless ggb_prof_enable(1) and not capturing
#define PROFILE(name, call)          \
    do {                             \
        ggb_prof_start(name);        \
        const int rc__ = (call);     \
        ggb_prof_stop(name);         \
        if (rc__) return rc__;       \
    } while (0)

int gemma_model::emit_decode_step() {
    const gemma_config &c = cfg_;
    const int64_t E = c.n_embd, F = c.n_ff, V = c.n_vocab;
    const int64_t QD = (int64_t)c.n_head * c.head_dim;
    const float q_scale = 1.0f / sqrtf((float)c.head_dim);                                          // :708
    PROFILE("op_get_rows", ggb_op_embed_token(token_embd_, state_, sqrtf((float)c.n_embd), x0_));   // :675-679
    const float *x_in = x0_;
    for (int il = 0; il < c.n_layer; il++) {
        const gemma_layer &l = layers_[(size_t)il];
        float *sa = sa_out_ + (size_t)il * E, *lo = l_out_ + (size_t)il * E, *kqv = kqv_ + (size_t)il * QD;
        PROFILE("op_rms_norm", ggb_op_rms_norm(x_in, l.attn_norm, c.eps, E, 1, nullptr, act_e_));   // :690
        const ggb_weight *qkv[3] = {l.attn_q, l.attn_k, l.attn_v};
        float *qkv_dst[3] = {q_, k_, v_};
        PROFILE("op_mul_mat", ggb_op_mul_mat_q8_multi(3, qkv, act_e_, qkv_dst));                    // :692-696
        PROFILE("op_rope", ggb_op_rope_kv_store(q_, k_, v_, state_ + 1, c.n_head, c.n_head_kv, c.head_dim, q_scale,
                                                k_cache_[(size_t)il], v_cache_[(size_t)il]));       // :698-716, :499-518
        PROFILE("op_attn", ggb_op_attn_decode(q_, k_cache_[(size_t)il], v_cache_[(size_t)il], state_ + 1, c.n_head,
                                              c.n_head_kv, c.head_dim, c.n_ctx, attn_scratch_, kqv, act_q_)); // :454-489
        PROFILE("op_mul_mat", ggb_op_mul_mat_q8(l.attn_output, act_q_, 1, sa, E, x_in));            // :493 + :723
        PROFILE("op_rms_norm", ggb_op_rms_norm(sa, l.ffn_norm, c.eps, E, 1, nullptr, act_e_));      // :724
        PROFILE("op_mul_mat", ggb_op_ffn_gate_up(l.ffn_gate, l.ffn_up, act_e_, act_f_,
                                                 capture_acts_ ? ffn_act_ + (size_t)il * F : nullptr)); // :446-449
        PROFILE("op_mul_mat", ggb_op_mul_mat_q8(l.ffn_down, act_f_, 1, lo, E, sa));                 // :450 + :731
        x_in = lo;
    }
    PROFILE("op_rms_norm", ggb_op_rms_norm(x_in, output_norm_, c.eps, E, 1, result_norm_, act_e_)); // :738
    PROFILE("op_mul_mat", ggb_op_mul_mat_q8(token_embd_, act_e_, 1, logits_, V, nullptr));          // :740 (tied)
    PROFILE("op_argmax", ggb_op_sample_greedy(logits_, V, state_, history_, sample_scratch_));      // :532-546
    return 0;
}

int gemma_model::kq_mv(const kq_mat &w, const float *x, float *dst) {
    // quantize_row_q8_K of the input row (ggml's INIT stage for a K-quant src0): inside the mat-vec's staging for short
    // rows (every CTA quantises the row itself: one launch less); as a launch of its own for long rows (the down
    // projection: 64 KB of fp32 per CTA would be read a thousand times)
    if (w.cols <= 4096) return ggb_op_mul_mat_kq_packed_f32(w.type, w.packed, w.rows, w.cols, x, 1, dst, w.rows);
    GM_TRY(ggb_op_quantize_q8_K(x, w.cols, 1, kq_q8_));
    const int64_t row_size = (int64_t)ggb_type_row_bytes(GGB_TYPE_Q8_K, w.cols);
    return w.type == GGB_TYPE_Q4_K ? ggb_op_mul_mat_q4k_packed(w.packed, w.rows, w.cols, kq_q8_, row_size, 1, dst, w.rows)
                                   : ggb_op_mul_mat_q6k_packed(w.packed, w.rows, w.cols, kq_q8_, row_size, 1, dst, w.rows);
}

// The same node list on K-quant weights: every mul_mat node = quantize_row_q8_K of its input (ggml's INIT stage for a
// K-quant src0) + the K-quant mat-vec; norms, rope, attention, GeGLU, residual adds and the sampler are the Q4_0 / Q8_0
// path's operators on fp32 rows.
int gemma_model::emit_decode_step_kq() {
    const gemma_config &c = cfg_;
    const int64_t E = c.n_embd, F = c.n_ff, V = c.n_vocab;
    const int64_t QD = (int64_t)c.n_head * c.head_dim;
    const float q_scale = 1.0f / sqrtf((float)c.head_dim);
    PROFILE("op_get_rows", ggb_op_get_rows_blocks(kq_embd_.type, kq_embd_.raw, V, (int64_t)ggb_type_row_bytes(kq_embd_.type, E), E,
                                                   state_, 1, sqrtf((float)c.n_embd), x0_));
    const float *x_in = x0_;
    for (int il = 0; il < c.n_layer; il++) {
        const gemma_layer &l = layers_[(size_t)il];
        const kq_mat *w = &kq_layers_[(size_t)il * 7];
        float *sa = sa_out_ + (size_t)il * E, *lo = l_out_ + (size_t)il * E, *kqv = kqv_ + (size_t)il * QD;
        float *act = ffn_act_ + (size_t)il * F;
        PROFILE("op_rms_norm", ggb_op_rms_norm(x_in, l.attn_norm, c.eps, E, 1, kq_xn_, nullptr));
        PROFILE("op_mul_mat", kq_mv(w[0], kq_xn_, q_));
        PROFILE("op_mul_mat", kq_mv(w[1], kq_xn_, k_));
        PROFILE("op_mul_mat", kq_mv(w[2], kq_xn_, v_));
        PROFILE("op_rope", ggb_op_rope_kv_store(q_, k_, v_, state_ + 1, c.n_head, c.n_head_kv, c.head_dim, q_scale,
                                                k_cache_[(size_t)il], v_cache_[(size_t)il]));
        PROFILE("op_attn", ggb_op_attn_decode(q_, k_cache_[(size_t)il], v_cache_[(size_t)il], state_ + 1, c.n_head,
                                              c.n_head_kv, c.head_dim, c.n_ctx, attn_scratch_, kqv, nullptr));
        PROFILE("op_mul_mat", kq_mv(w[3], kqv, sa));
        PROFILE("op_add", ggb_op_add(sa, x_in, E, sa));
        PROFILE("op_rms_norm", ggb_op_rms_norm(sa, l.ffn_norm, c.eps, E, 1, kq_xn_, nullptr));
        PROFILE("op_mul_mat", kq_mv(w[4], kq_xn_, kq_gate_));
        PROFILE("op_mul_mat", kq_mv(w[5], kq_xn_, kq_up_));
        PROFILE("op_gelu", ggb_op_gelu(kq_gate_, F, kq_gate_));
        PROFILE("op_mul", ggb_op_mul(kq_gate_, kq_up_, F, act));
        PROFILE("op_mul_mat", kq_mv(w[6], act, lo));
        PROFILE("op_add", ggb_op_add(lo, sa, E, lo));
        x_in = lo;
    }
    PROFILE("op_rms_norm", ggb_op_rms_norm(x_in, output_norm_, c.eps, E, 1, result_norm_, nullptr));
    PROFILE("op_mul_mat", kq_mv(kq_embd_, result_norm_, logits_));
    PROFILE("op_argmax", ggb_op_sample_greedy(logits_, V, state_, history_, sample_scratch_));
    return 0;
}

// The same token step, row-sharded over plan_.size ranks.  Every rank keeps the replicated
// fp32 residual stream; each weight mat-vec computes this rank's rows only and the shards meet in four
// allgathers per layer: attention output, o-proj (+residual), GeGLU activations (already Q8_0: int8 codes and
// block scales are gathered, 4.25x fewer bytes than fp32), down-proj (+residual); plus 8 bytes per rank for the
// arg-max.  Collectives are NCCL on the compute stream, captured in the same CUDA graph.
int gemma_model::emit_decode_step_tp() {
    const gemma_config &c = cfg_;
    const tp_plan &p = plan_;
    const int64_t E = c.n_embd, F = c.n_ff, V = c.n_vocab, d = c.head_dim;
    const int64_t QD = (int64_t)c.n_head * d;
    const float q_scale = 1.0f / sqrtf((float)d);
    PROFILE("op_get_rows", ggb_op_embed_token(token_embd_, state_, sqrtf((float)c.n_embd), x0_));
    const float *x_in = x0_;
    uint8_t *actf = (uint8_t *)act_f_;
    for (int il = 0; il < c.n_layer; il++) {
        const gemma_layer &l = layers_[(size_t)il];
        float *sa = sa_out_ + (size_t)il * E, *lo = l_out_ + (size_t)il * E, *kqv = kqv_ + (size_t)il * QD;
        PROFILE("op_rms_norm", ggb_op_rms_norm(x_in, l.attn_norm, c.eps, E, 1, nullptr, act_e_));
        const ggb_weight *qkv[3] = {l.sh_q, l.sh_k, l.sh_v};
        float *qkv_dst[3] = {q_, k_, v_};
        PROFILE("op_mul_mat", ggb_op_mul_mat_q8_multi(3, qkv, act_e_, qkv_dst));
        PROFILE("op_rope", ggb_op_rope_kv_store(q_, k_, v_, state_ + 1, p.heads, p.kv_heads, c.head_dim, q_scale,
                                                k_cache_[(size_t)il], v_cache_[(size_t)il]));
        PROFILE("op_attn", ggb_op_attn_decode(q_, k_cache_[(size_t)il], v_cache_[(size_t)il], state_ + 1, p.heads,
                                              p.kv_heads, c.head_dim, c.n_ctx, attn_scratch_, kqv + p.q.row0, nullptr));
        PROFILE("op_allgather", ggb_op_allgather(kqv + p.q.row0, kqv, (size_t)p.q.nrows * 4));
        PROFILE("op_quantize", ggb_op_quantize_q8_0(kqv, QD, 1, act_q_));
        PROFILE("op_mul_mat", ggb_op_mul_mat_q8(l.sh_o, act_q_, 1, sa + p.o.row0, E, x_in + p.o.row0));
        PROFILE("op_allgather", ggb_op_allgather(sa + p.o.row0, sa, (size_t)p.o.nrows * 4));
        PROFILE("op_rms_norm", ggb_op_rms_norm(sa, l.ffn_norm, c.eps, E, 1, nullptr, act_e_));
        PROFILE("op_mul_mat", ggb_op_ffn_gate_up_shard(l.sh_gate, l.sh_up, act_e_, actf + p.gate_up.row0,
                                                       actf + F + p.gate_up.row0 / 32 * 2,
                                                       capture_acts_ ? ffn_act_ + (size_t)il * F + p.gate_up.row0 : nullptr));
        PROFILE("op_allgather", ggb_op_allgather(actf + p.gate_up.row0, actf, (size_t)p.gate_up.nrows));
        PROFILE("op_allgather", ggb_op_allgather(actf + F + p.gate_up.row0 / 32 * 2, actf + F, (size_t)p.gate_up.nrows / 32 * 2));
        if (capture_acts_)
            PROFILE("op_allgather", ggb_op_allgather(ffn_act_ + (size_t)il * F + p.gate_up.row0, ffn_act_ + (size_t)il * F,
                                                     (size_t)p.gate_up.nrows * 4));
        PROFILE("op_mul_mat", ggb_op_mul_mat_q8(l.sh_down, act_f_, 1, lo + p.down.row0, E, sa + p.down.row0));
        PROFILE("op_allgather", ggb_op_allgather(lo + p.down.row0, lo, (size_t)p.down.nrows * 4));
        x_in = lo;
    }
    PROFILE("op_rms_norm", ggb_op_rms_norm(x_in, output_norm_, c.eps, E, 1, result_norm_, act_e_));
    PROFILE("op_mul_mat", ggb_op_mul_mat_q8(sh_logits_, act_e_, 1, logits_ + p.logits.row0, V, nullptr));
    PROFILE("op_argmax", ggb_op_argmax_shard(logits_ + p.logits.row0, p.logits.nrows, (int32_t)p.logits.row0,
                                             (uint8_t *)cand_ + 8 * p.rank, sample_scratch_));
    PROFILE("op_allgather", ggb_op_allgather((uint8_t *)cand_ + 8 * p.rank, cand_, 8));
    PROFILE("op_argmax", ggb_op_sample_from_candidates(cand_, p.size, state_, history_));
    return 0;
}

// eager (un-captured) steps with the cudaEvent probes live: per-op-class device time
int gemma_model::run_steps_profiled(int steps) {
    if (!built_) return GGB_ERR_ARG;
    ggb_prof_enable(1);
    int rc = 0;
    for (int i = 0; i < steps && !rc; i++) rc = kq_mode_ ? emit_decode_step_kq() : plan_.size > 1 ? emit_decode_step_tp() : emit_decode_step();
    ggb_sync();
    ggb_prof_enable(0);
    return rc;
}

// ref: build_compute_graph (:665-747).  fused: the whole token step is the persistent decode kernel;
// otherwise the ggb_op_* sequence of emit_decode_step().  Either way the step is captured into a
// CUDA graph once and replayed per token (a cooperative launch that cannot be captured is
// launched directly).
int gemma_model::build_compute_graph(bool capture_acts, bool fused) {
    if (kq_mode_) {
        if (!kq_embd_.packed || !output_norm_ || k_cache_.empty() || plan_.size > 1) return GGB_ERR_ARG;
        for (const kq_mat &w : kq_layers_)
            if (!w.packed) return GGB_ERR_ARG;
        for (const gemma_layer &l : layers_)
            if (!l.attn_norm || !l.ffn_norm) return GGB_ERR_ARG;
        fused = false; // the persistent kernel streams Q4_0 / Q8_0 only
    } else if (!token_embd_ || k_cache_.empty()) {
        return GGB_ERR_ARG;
    }
    GM_TRY(alloc_buffers());
    if (kq_mode_ && !kq_q8_) {
        const gemma_config &c = cfg_;
        const int64_t kmax = std::max<int64_t>(std::max<int64_t>(c.n_embd, c.n_ff), (int64_t)c.n_head * c.head_dim);
        GM_TRY(ggb_buffer_alloc((size_t)c.n_embd * 4, (void **)&kq_xn_));
        GM_TRY(ggb_buffer_alloc((size_t)c.n_ff * 4, (void **)&kq_gate_));
        GM_TRY(ggb_buffer_alloc((size_t)c.n_ff * 4, (void **)&kq_up_));
        GM_TRY(ggb_buffer_alloc(ggb_type_row_bytes(GGB_TYPE_Q8_K, kmax), &kq_q8_));
    }
    capture_acts_ = capture_acts;
    ggb_graph_free(graph_);
    graph_ = nullptr;
    ggb_decoder_free(decoder_);
    decoder_ = nullptr;
    built_ = false;
    GM_TRY(ggb_sync());
    if (!kq_mode_) GM_TRY(make_shards());
    if (plan_.size > 1) fused = tp_fused_; // multi-GPU: persistent kernel with peer stores, or ops + NCCL
    if (fused) {
        const gemma_config &c = cfg_;
        ggb_decoder_dims dims{c.n_embd, c.n_layer, c.n_head, c.n_head_kv, c.head_dim, c.n_ff, c.n_vocab, c.n_ctx, c.eps, 0};
        std::vector<ggb_decoder_layer> ls((size_t)c.n_layer);
        for (int il = 0; il < c.n_layer; il++) {
            const gemma_layer &l = layers_[(size_t)il];
            ls[(size_t)il] = ggb_decoder_layer{l.attn_norm, l.ffn_norm, l.attn_q, l.attn_k, l.attn_v, l.attn_output,
                                               l.ffn_gate, l.ffn_up, l.ffn_down, k_cache_[(size_t)il], v_cache_[(size_t)il]};
        }
        ggb_decoder_io io{state_, history_, sa_out_, l_out_, kqv_, capture_acts ? ffn_act_ : nullptr, result_norm_, logits_};
        const int rc = ggb_decoder_create_tp(&dims, ls.data(), token_embd_, output_norm_, &io, plan_.size > 1 ? plan_.rank : 0,
                                             plan_.size > 1 ? plan_.size : 1, &decoder_);
        if (rc == GGB_ERR_UNSUPPORTED && plan_.size > 1) return rc; // the KV cache is already in the fused layout
        if (rc == GGB_ERR_UNSUPPORTED) {
            decoder_ = nullptr; // shape outside the fused envelope: lower node by node instead
        } else if (rc) {
            return rc;
        }
    }
    const int rcb = ggb_capture_begin();
    if (rcb == 0) {
        const int rc = decoder_ ? ggb_decoder_step(decoder_)
                                : (kq_mode_ ? emit_decode_step_kq() : plan_.size > 1 ? emit_decode_step_tp() : emit_decode_step());
        ggb_graph *g = nullptr;
        const int rc2 = ggb_capture_end(&g);
        if (rc == 0 && rc2 == 0) {
            graph_ = g;
        } else {
            ggb_graph_free(g);
            if (!decoder_) return rc ? rc : rc2; // the op sequence must be capturable
        }
    } else if (!decoder_) {
        return rcb; // no graph and no persistent kernel: there would be nothing to launch
    }
    built_ = true;
    return 0;
}

int gemma_model::launch_step() {
    if (graph_) return ggb_graph_launch(graph_);
    if (decoder_) return ggb_decoder_step(decoder_);
    return GGB_ERR_ARG;
}

int gemma_model::set_state(token_id token, int32_t pos, int32_t pos_step) {
    if (pos < 0 || pos >= cfg_.n_ctx) return GGB_ERR_ARG;
    h_state_[0] = token;
    h_state_[1] = pos;
    h_state_[2] = pos_step;
    h_state_[3] = cfg_.n_ctx; // the device stops advancing the position at n_ctx - 1
    pos_hint_ = pos;
    pos_step_hint_ = pos_step;
    return ggb_upload(state_, h_state_, 16);
}

int gemma_model::run_steps(int steps) {
    if (!built_ || steps < 0) return GGB_ERR_ARG;
    // stepping past the context would address the KV cache / rope table / history out of bounds: refuse
    if (pos_step_hint_ > 0 && pos_hint_ + (int64_t)(steps - 1) * pos_step_hint_ > (int64_t)cfg_.n_ctx - 1) return GGB_ERR_ARG;
    for (int i = 0; i < steps; i++) GM_TRY(launch_step());
    pos_hint_ += (int64_t)steps * pos_step_hint_;
    return 0;
}

int gemma_model::get_state(token_id *token, int32_t *pos) {
    int32_t tmp[4];
    GM_TRY(ggb_download(tmp, state_, 16));
    if (token) *token = tmp[0];
    if (pos) *pos = tmp[1];
    return 0;
}

// host-buffer step: the e2e path (pinned H2D of the inputs, graph, D2H of the sampled token)
int gemma_model::decode_step(token_id token, int32_t pos, token_id *next) {
    GM_TRY(set_state(token, pos, 1));
    GM_TRY(launch_step());
    GM_TRY(ggb_download(h_state_ + 4, state_, 4)); // pinned: a true async D2H + stream sync
    if (next) *next = h_state_[4];
    return 0;
}

// ref: inference() (:231-286).  PREFILL walks the prompt one position at a time through the decode
// graph (bit-identical to the reference's batched prefill because every column of a ggml mul_mat and
// every soft-max row is computed independently); the tcgen05 GEMM prefill is a separate path.
int gemma_model::inference(std::vector<token_id> &input, inference_stage stage) {
    if (input.empty() || !built_) return GGB_ERR_ARG;
    token_id next = -1;
    if (stage == inference_stage::PREFILL) {
        for (size_t t = 0; t < input.size(); t++) GM_TRY(decode_step(input[t], (int32_t)t, &next));
    } else {
        GM_TRY(decode_step(input.back(), (int32_t)input.size() - 1, &next));
    }
    input.push_back(next);
    return 0;
}

// ref: begin_one_round_inference (:548-575) with a caller-supplied prompt and token budget
int gemma_model::generate(const std::vector<token_id> &prompt, int n_decode, std::vector<token_id> &out,
                          double *prefill_ms, double *decode_ms) {
    using clk = std::chrono::high_resolution_clock;
    std::vector<token_id> input = prompt;
    GM_TRY(init_kv_cache());
    const auto t0 = clk::now();
    GM_TRY(inference(input, inference_stage::PREFILL));
    const auto t1 = clk::now();
    while ((int)(input.size() - prompt.size()) < n_decode && (int)input.size() < cfg_.n_ctx)
        GM_TRY(inference(input, inference_stage::DECODE));
    const auto t2 = clk::now();
    out.assign(input.begin() + (long)prompt.size(), input.end());
    if (prefill_ms) *prefill_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (decode_ms) *decode_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    return 0;
}

// ---------------------------------------------------------------- batched prefill (tcgen05 GEMMs)
void gemma_model::free_prefill() {
    void *bufs[] = {pf_ids_, pf_pos_, pf_x_, pf_xn_, pf_q_, pf_k_, pf_v_, pf_attn_, pf_sa_, pf_g_, pf_u_, pf_logits_, pf_a_, pf_ws_};
    for (void *b : bufs)
        if (b) ggb_buffer_free(b);
    pf_ids_ = pf_pos_ = nullptr;
    pf_x_ = pf_xn_ = pf_q_ = pf_k_ = pf_v_ = pf_attn_ = pf_sa_ = pf_g_ = pf_u_ = pf_logits_ = nullptr;
    pf_a_ = pf_ws_ = nullptr;
    pf_T_ = 0;
}

int gemma_model::alloc_prefill(int T) {
    if (T <= pf_T_) return 0;
    free_prefill();
    const gemma_config &c = cfg_;
    const size_t E = (size_t)c.n_embd, F = (size_t)c.n_ff, V = (size_t)c.n_vocab, t = (size_t)T;
    const size_t QD = (size_t)c.n_head * c.head_dim, KD = (size_t)c.n_head_kv * c.head_dim;
    size_t widest = E > QD ? E : QD;
    if (F > widest) widest = F;
    GM_TRY(ggb_buffer_alloc(t * 4, (void **)&pf_ids_));
    GM_TRY(ggb_buffer_alloc(t * 4, (void **)&pf_pos_));
    GM_TRY(ggb_buffer_alloc(t * E * 4, (void **)&pf_x_));
    GM_TRY(ggb_buffer_alloc(t * E * 4, (void **)&pf_xn_));
    GM_TRY(ggb_buffer_alloc(t * QD * 4, (void **)&pf_q_));
    GM_TRY(ggb_buffer_alloc(t * KD * 4, (void **)&pf_k_));
    GM_TRY(ggb_buffer_alloc(t * KD * 4, (void **)&pf_v_));
    GM_TRY(ggb_buffer_alloc(t * QD * 4, (void **)&pf_attn_));
    GM_TRY(ggb_buffer_alloc(t * E * 4, (void **)&pf_sa_));
    GM_TRY(ggb_buffer_alloc(t * F * 4, (void **)&pf_g_));
    GM_TRY(ggb_buffer_alloc(t * F * 4, (void **)&pf_u_));
    GM_TRY(ggb_buffer_alloc(t * V * 4, (void **)&pf_logits_));
    GM_TRY(ggb_buffer_alloc(t * widest * 2, &pf_a_));
    // the K-split is only taken for small N (o-proj, down-proj, the fused q | k | v launch)
    const size_t KDp = (size_t)cfg_.n_head_kv * cfg_.head_dim;
    const size_t qkv_n = (QD + 255) / 256 * 256 + (KDp + 255) / 256 * 256 + KDp;
    pf_ws_n_ = (int64_t)std::max(std::max(E, QD), qkv_n);
    pf_ws_bytes_ = ggb_gemm_workspace_bytes((int64_t)t, pf_ws_n_, 0);
    GM_TRY(ggb_buffer_alloc(pf_ws_bytes_, &pf_ws_));
    pf_T_ = T;
    return 0;
}

int gemma_model::prefill(const token_id *tokens, int T, int32_t pos0, token_id *next) {
    const gemma_config &c = cfg_;
    if (!tokens || T <= 0 || pos0 < 0 || pos0 + T > c.n_ctx || !built_ || plan_.size > 1) return GGB_ERR_ARG;
    if (kq_mode_) return GGB_ERR_UNSUPPORTED; // the GEMM prefill dequantises Q4_0 / Q8_0; a K-quant model prefills token by token
    GM_TRY(alloc_prefill(T));
    const int64_t E = c.n_embd, F = c.n_ff, V = c.n_vocab, t = T;
    const int64_t QD = (int64_t)c.n_head * c.head_dim;
    const float q_scale = 1.0f / sqrtf((float)c.head_dim);
    std::vector<int32_t> pos((size_t)T);
    for (int i = 0; i < T; i++) pos[(size_t)i] = pos0 + i;
    GM_TRY(ggb_upload(pf_ids_, tokens, (size_t)T * 4));
    GM_TRY(ggb_upload(pf_pos_, pos.data(), (size_t)T * 4));
    auto gemm = [&](const ggb_weight *w, int64_t K, float *out, const float *resid) -> int {
        pf_flops_ += 2.0 * (double)t * (double)K * (double)ggb_weight_rows(w);
        const int64_t N = ggb_weight_rows(w);
        const bool small = N <= pf_ws_n_;
        return ggb_op_gemm_bf16(w, pf_a_, t, out, N, resid, small ? pf_ws_ : nullptr, small ? pf_ws_bytes_ : 0);
    };
    // matrices that share the A operand go out as ONE launch (more tiles per launch, one pass over A)
    auto gemm_multi = [&](int n, const ggb_weight *const *w, int64_t K, float *const *outs) -> int {
        int64_t ld[3], total = 0;
        for (int i = 0; i < n; i++) {
            ld[i] = ggb_weight_rows(w[i]);
            pf_flops_ += 2.0 * (double)t * (double)K * (double)ld[i];
            total += i + 1 < n ? (ld[i] + 255) / 256 * 256 : ld[i];
        }
        const bool small = total <= pf_ws_n_;
        return ggb_op_gemm_bf16_multi(n, w, pf_a_, t, outs, ld, small ? pf_ws_ : nullptr, small ? pf_ws_bytes_ : 0);
    };
    pf_flops_ = 0.0;
    PROFILE("op_get_rows", ggb_op_get_rows(token_embd_, pf_ids_, t, sqrtf((float)c.n_embd), pf_x_));   // :675-679
    for (int il = 0; il < c.n_layer; il++) {
        const gemma_layer &l = layers_[(size_t)il];
        void *kc = k_cache_[(size_t)il], *vc = v_cache_[(size_t)il];
        PROFILE("op_rms_norm", ggb_op_rms_norm_bf16(pf_x_, l.attn_norm, c.eps, E, t, pf_a_));           // :690
        {
            const ggb_weight *w3[3] = {l.attn_q, l.attn_k, l.attn_v};
            float *o3[3] = {pf_q_, pf_k_, pf_v_};
            PROFILE("op_mul_mat", gemm_multi(3, w3, E, o3));                                             // :692-696
        }
        // rope(q) * q_scale in place, rope(k) and v straight into the position-major fp16 cache rows pos0 .. pos0+T-1
        // (ref :698-716 rope + scale, :499-518 KV store): one launch
        PROFILE("op_rope", ggb_op_rope_kv_store_batch(pf_q_, pf_k_, pf_v_, pf_pos_, t, c.n_head, c.n_head_kv, c.head_dim,
                                                       q_scale, kc, vc));
        // kqv goes straight out as the bf16 A operand of the o-projection (the fp32 copy only for captures)
        PROFILE("op_attn", ggb_op_attn_prefill(pf_q_, kc, vc, t, pos0, c.n_head, c.n_head_kv, c.head_dim,
                                               capture_acts_ ? pf_attn_ : nullptr, pf_a_)); // :454-489
        PROFILE("op_mul_mat", gemm(l.attn_output, QD, pf_sa_, pf_x_));                                   // :493 + :723
        PROFILE("op_rms_norm", ggb_op_rms_norm_bf16(pf_sa_, l.ffn_norm, c.eps, E, t, pf_a_));            // :727
        {
            const ggb_weight *w2[2] = {l.ffn_gate, l.ffn_up};
            float *o2[2] = {pf_g_, pf_u_};
            PROFILE("op_mul_mat", gemm_multi(2, w2, E, o2));                                             // :446-447
        }
        PROFILE("op_gelu", ggb_op_geglu_bf16(pf_g_, pf_u_, t * F, pf_a_));                               // :448-449
        PROFILE("op_mul_mat", gemm(l.ffn_down, F, pf_x_, pf_sa_));                                       // :450 + :731
    }
    PROFILE("op_rms_norm", ggb_op_rms_norm_bf16(pf_x_, output_norm_, c.eps, E, t, pf_a_));               // :735
    PROFILE("op_mul_mat", gemm(token_embd_, E, pf_logits_, nullptr));                                    // :740 (all T rows)
    // greedy token of the last position; decode continues at pos0 + T
    h_state_[0] = 0; h_state_[1] = pos0 + T - 1; h_state_[2] = 1; h_state_[3] = c.n_ctx;
    pos_hint_ = pos0 + T; pos_step_hint_ = 1;
    GM_TRY(ggb_upload(state_, h_state_, 16));
    PROFILE("op_argmax", ggb_op_sample_greedy(pf_logits_ + (size_t)(T - 1) * (size_t)V, V, state_, history_, sample_scratch_));
    GM_TRY(ggb_download(h_state_ + 4, state_, 16));
    if (next) *next = h_state_[4];
    pf_last_T_ = T;
    return 0;
}

int gemma_model::get_prefill_logits(int row, float *host) const {
    if (!pf_logits_ || row < 0 || row >= pf_last_T_) return GGB_ERR_ARG;
    return ggb_download(host, pf_logits_ + (size_t)row * (size_t)cfg_.n_vocab, (size_t)cfg_.n_vocab * 4);
}

int gemma_model::get_logits(float *host) const {
    if (plan_.size > 1) { // the shards live at their own offsets of logits_: gather them (collective: every rank calls)
        GM_TRY(ggb_op_allgather(logits_ + plan_.logits.row0, logits_, (size_t)plan_.logits.nrows * 4));
    }
    return ggb_download(host, logits_, (size_t)cfg_.n_vocab * 4);
}

// tensor_dump-style access to named intermediates of the last step (ref: src/tensor_dump.cpp:11-27)
int gemma_model::get_capture(const std::string &name, int layer, float *host, int64_t n) const {
    const size_t E = (size_t)cfg_.n_embd, F = (size_t)cfg_.n_ff, QD = (size_t)cfg_.n_head * cfg_.head_dim;
    const float *src = nullptr;
    size_t len = 0;
    if (layer < 0 || layer >= cfg_.n_layer) layer = 0;
    if (name == "l_out") { src = l_out_ + (size_t)layer * E; len = E; }
    else if (name == "sa_out") { src = sa_out_ + (size_t)layer * E; len = E; }
    else if (name == "kqv_out") { src = kqv_ + (size_t)layer * QD; len = QD; }
    else if (name == "ffn_act") { src = capture_acts_ ? ffn_act_ + (size_t)layer * F : nullptr; len = F; }
    else if (name == "result_norm") { src = result_norm_; len = E; }
    else if (name == "result_output") { src = logits_; len = (size_t)cfg_.n_vocab; }
    if (!src || (size_t)n != len) return GGB_ERR_ARG;
    return ggb_download(host, src, len * 4);
}

int gemma_model::get_kv(int layer, void *k_host, void *v_host) const {
    if (layer < 0 || layer >= (int)k_cache_.size()) return GGB_ERR_ARG;
    const size_t bytes = (size_t)cfg_.n_ctx * kv_heads_alloc() * cfg_.head_dim * 2;
    if (k_host) GM_TRY(ggb_download(k_host, k_cache_[(size_t)layer], bytes));
    if (v_host) GM_TRY(ggb_download(v_host, v_cache_[(size_t)layer], bytes));
    return 0;
}
