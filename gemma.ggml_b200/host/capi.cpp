// capi.cpp — flat C entry points over gemma_model so tests / bench.py can drive the host side with
// ctypes (the reference drives its model from app::run, ref: src/app.cpp:30-44).
#include <cstring>
#include <string>
#include <vector>

#include "gemma_model.h"
#include "gguf_loader.h"
#include "tensor_dump.h"

extern "C" {

void *gm_create(const gemma_config *cfg) { return new gemma_model(*cfg); }
void gm_free(void *m) { delete (gemma_model *)m; }
int gm_init_tensor_parallel(void *m, int rank, int size) { return ((gemma_model *)m)->init_tensor_parallel(rank, size); }
// host-only: fills out[18] = {valid, head0, heads, kv_head0, kv_heads, kv_replicated, q.row0, q.nrows, k.row0, k.nrows,
// o.row0, o.nrows, gate_up.row0, gate_up.nrows, down.row0, down.nrows, logits.row0, logits.nrows}
int gm_tp_plan(const gemma_config *cfg, int rank, int size, int64_t *out) {
    const tp_plan p = make_tp_plan(*cfg, rank, size);
    const int64_t v[18] = {p.valid, p.head0, p.heads, p.kv_head0, p.kv_heads, p.kv_replicated, p.q.row0, p.q.nrows,
                           p.k.row0, p.k.nrows, p.o.row0, p.o.nrows, p.gate_up.row0, p.gate_up.nrows, p.down.row0,
                           p.down.nrows, p.logits.row0, p.logits.nrows};
    for (int i = 0; i < 18; i++) out[i] = v[i];
    return p.valid ? 0 : -2;
}
int gm_init_synthetic_weights(void *m) { return ((gemma_model *)m)->init_synthetic_weights(); }
int gm_set_tensor(void *m, int tensor_id, const void *host, int64_t nbytes) {
    return ((gemma_model *)m)->set_tensor(tensor_id, host, nbytes);
}
int gm_set_tensor_typed(void *m, int tensor_id, int type, const void *host, int64_t nbytes) {
    return ((gemma_model *)m)->set_tensor_typed(tensor_id, type, host, nbytes);
}
int gm_kquant_mode(void *m) { return ((gemma_model *)m)->kquant_mode() ? 1 : 0; }
int gm_init_kv_cache(void *m) { return ((gemma_model *)m)->init_kv_cache(); }
int gm_fill_kv_cache_synthetic(void *m, uint64_t seed) { return ((gemma_model *)m)->fill_kv_cache_synthetic(seed); }
int gm_build_compute_graph(void *m, int capture_acts, int fused) {
    return ((gemma_model *)m)->build_compute_graph(capture_acts != 0, fused != 0);
}
void *gm_decoder(void *m) { return ((gemma_model *)m)->decoder(); }
int gm_is_fused(void *m) { return ((gemma_model *)m)->is_fused() ? 1 : 0; }
int gm_decode_step(void *m, int32_t token, int32_t pos, int32_t *next) {
    return ((gemma_model *)m)->decode_step(token, pos, next);
}
int gm_set_state(void *m, int32_t token, int32_t pos, int32_t pos_step) {
    return ((gemma_model *)m)->set_state(token, pos, pos_step);
}
int gm_run_steps(void *m, int steps) { return ((gemma_model *)m)->run_steps(steps); }
int gm_run_steps_profiled(void *m, int steps) { return ((gemma_model *)m)->run_steps_profiled(steps); }
int gm_get_state(void *m, int32_t *token, int32_t *pos) { return ((gemma_model *)m)->get_state(token, pos); }
int gm_generate(void *m, const int32_t *prompt, int n_prompt, int n_decode, int32_t *out, int *n_out, double *prefill_ms,
                double *decode_ms) {
    std::vector<token_id> p(prompt, prompt + n_prompt), o;
    const int rc = ((gemma_model *)m)->generate(p, n_decode, o, prefill_ms, decode_ms);
    if (rc) return rc;
    if (n_out) *n_out = (int)o.size();
    memcpy(out, o.data(), o.size() * sizeof(int32_t));
    return 0;
}
// ---- GGUF loader + tokenizer.  gm_gguf_probe parses a file on the host only (no GPU needed).
static thread_local std::string g_capi_err;
const char *gm_last_error() { return g_capi_err.c_str(); }
int gm_gguf_probe(const char *path, gemma_config *cfg, int32_t *n_tensors, int32_t *n_kv, int32_t *n_tokens, int32_t *bos, int32_t *eos) {
    gguf_file f;
    if (f.open(path) != 0) { g_capi_err = f.error(); return -1; }
    if (gguf_gemma_config(f, 0, cfg, &g_capi_err) != 0) return -2;
    gemma_tokenizer tok;
    const bool has_tok = gguf_gemma_tokenizer(f, &tok) == 0;
    if (n_tensors) *n_tensors = (int32_t)f.tensors().size();
    if (n_kv) *n_kv = (int32_t)f.n_kv();
    if (n_tokens) *n_tokens = has_tok ? (int32_t)tok.tokens.size() : 0;
    if (bos) *bos = tok.special_bos_id;
    if (eos) *eos = tok.special_eos_id;
    return 0;
}
// detokenise with the vocabulary of a GGUF file (host only): "▁" -> " ", leading "<bos>" dropped; returns the length
int gm_gguf_detokenize(const char *path, const int32_t *ids, int n, char *buf, int cap) {
    gguf_file f;
    if (f.open(path) != 0) { g_capi_err = f.error(); return -1; }
    gemma_tokenizer tok;
    if (gguf_gemma_tokenizer(f, &tok) != 0) { g_capi_err = "no tokenizer.ggml.tokens in the file"; return -2; }
    const std::string t = tok.tokens_text(std::vector<token_id>(ids, ids + n));
    if (buf && cap > 0) {
        const size_t k = t.size() < (size_t)cap - 1 ? t.size() : (size_t)cap - 1;
        memcpy(buf, t.data(), k);
        buf[k] = 0;
    }
    return (int)t.size();
}
void *gm_load_model_from_file(const char *path, int32_t n_ctx) {
    gemma_model *m = gemma_model::load_model_from_file(path, n_ctx, &g_capi_err);
    return m;
}
int gm_get_config(void *m, gemma_config *cfg) {
    *cfg = ((gemma_model *)m)->config();
    return 0;
}
int gm_begin_one_round_inference(void *m, const int32_t *prompt, int n_prompt, int max_tokens, int32_t *out, int *n_out,
                                 char *text, int text_cap, double *prefill_ms, double *decode_ms) {
    std::vector<token_id> p(prompt, prompt + n_prompt), o;
    std::string t;
    const int rc = ((gemma_model *)m)->begin_one_round_inference(p, max_tokens, o, text ? &t : nullptr, prefill_ms, decode_ms);
    if (rc) return rc;
    if (n_out) *n_out = (int)o.size();
    memcpy(out, o.data(), o.size() * sizeof(int32_t));
    if (text && text_cap > 0) {
        const size_t k = t.size() < (size_t)text_cap - 1 ? t.size() : (size_t)text_cap - 1;
        memcpy(text, t.data(), k);
        text[k] = 0;
    }
    return 0;
}
int gm_prefill(void *m, const int32_t *tokens, int T, int32_t pos0, int32_t *next) {
    return ((gemma_model *)m)->prefill(tokens, T, pos0, next);
}
int gm_get_prefill_logits(void *m, int row, float *host) { return ((gemma_model *)m)->get_prefill_logits(row, host); }
double gm_last_prefill_gemm_flops(void *m) { return ((gemma_model *)m)->last_prefill_gemm_flops(); }
int gm_get_logits(void *m, float *host) { return ((gemma_model *)m)->get_logits(host); }
int gm_get_capture(void *m, const char *name, int layer, float *host, int64_t n) {
    return ((gemma_model *)m)->get_capture(name, layer, host, n);
}
int gm_get_kv(void *m, int layer, void *k_host, void *v_host) { return ((gemma_model *)m)->get_kv(layer, k_host, v_host); }
int gm_kernels_per_step(void *m) { return ((gemma_model *)m)->kernels_per_step(); }
uint64_t gm_weight_bytes(void *m) { return (uint64_t)((gemma_model *)m)->weight_bytes(); }

// ---- tensor_dump (ref: src/tensor_dump.h): host-side parity harness
void td_set_dir(const char *dir) { set_tensor_dump_dir(dir); }
int td_dump_ptr_data(const char *name, const void *ptr, uint64_t size, int target) {
    return dump_ptr_data(name, ptr, (size_t)size, target ? tensor_dump_mode::target : tensor_dump_mode::source) ? 0 : -1;
}
// a named intermediate of the model's last step, fetched from the device and dumped as "<name>-<layer>_source|_target"
int td_dump_capture(void *m, const char *name, int layer, int64_t n, int target) {
    std::vector<float> h((size_t)n);
    const int rc = ((gemma_model *)m)->get_capture(name, layer, h.data(), n);
    if (rc) return rc;
    const std::string file = std::string(name) + "-" + std::to_string(layer);
    return dump_ptr_data(file.c_str(), h.data(), h.size() * 4, target ? tensor_dump_mode::target : tensor_dump_mode::source) ? 0 : -1;
}
int td_compare_tensors(const char *name) { return compare_tensors(name) ? 1 : 0; }
// out[5] = {n, n_diff, max_abs, max_rel, max_target}; returns 0 if both files exist with equal fp32 sizes
int td_compare_tensors_report(const char *name, double *out) {
    tensor_compare_report r;
    if (!compare_tensors_report(name, &r)) return -1;
    out[0] = (double)r.n; out[1] = (double)r.n_diff; out[2] = r.max_abs; out[3] = r.max_rel; out[4] = r.max_target;
    return 0;
}

} // extern "C"
