// gguf_loader.cpp — see gguf_loader.h.  GGUF v2 / v3 layout: magic "GGUF", u32 version, u64 tensor count, u64 kv count;
// kv = string key, u32 type, value; tensor info = string name, u32 n_dims, u64 dims[], u32 ggml type, u64 offset; the
// data section starts at the next multiple of general.alignment (default 32).
#include "gguf_loader.h"

#include <cstdio>
#include <cstring>

namespace {

enum { T_U8, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64 };

struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    template <typename T> T get() {
        T v{};
        if (p + sizeof(T) > end) { ok = false; return v; }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        const uint64_t n = get<uint64_t>();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return {}; }
        std::string s((const char *)p, (size_t)n);
        p += n;
        return s;
    }
};

size_t scalar_size(uint32_t t) {
    switch (t) {
    case T_U8: case T_I8: case T_BOOL: return 1;
    case T_U16: case T_I16: return 2;
    case T_U32: case T_I32: case T_F32: return 4;
    case T_U64: case T_I64: case T_F64: return 8;
    default: return 0;
    }
}

// bytes of a tensor of ggml type `type` with `n` elements (the types this path knows)
uint64_t type_bytes(int32_t type, uint64_t n) {
    switch (type) {
    case 0: return n * 4;                // F32
    case 1: return n * 2;                // F16
    case 2: return n / 32 * 18;          // Q4_0
    case 8: return n / 32 * 34;          // Q8_0
    case 12: return n / 256 * 144;       // Q4_K
    case 14: return n / 256 * 210;       // Q6_K
    default: return 0;
    }
}

} // namespace

int gguf_file::open(const char *path) {
    FILE *fp = path ? fopen(path, "rb") : nullptr;
    if (!fp) { err_ = std::string("cannot open ") + (path ? path : "(null)"); return -1; }
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    blob_.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = blob_.empty() ? 0 : fread(blob_.data(), 1, blob_.size(), fp);
    fclose(fp);
    if (got != blob_.size() || blob_.size() < 24) { err_ = "short read"; return -1; }
    Reader r{blob_.data(), blob_.data() + blob_.size()};
    if (r.get<uint32_t>() != 0x46554747u) { err_ = "not a GGUF file (magic)"; return -2; }
    version_ = r.get<uint32_t>();
    if (version_ < 2 || version_ > 3) { err_ = "unsupported GGUF version " + std::to_string(version_); return -2; }
    const uint64_t n_tensors = r.get<uint64_t>(), n_kv = r.get<uint64_t>();
    // every key / tensor entry takes at least 12 bytes: counts beyond that are a corrupt header, not a reason to allocate
    if (n_tensors > blob_.size() / 12 || n_kv > blob_.size() / 12) { err_ = "corrupt GGUF header (counts)"; return -3; }
    for (uint64_t i = 0; i < n_kv && r.ok; i++) {
        const std::string key = r.str();
        const uint32_t type = r.get<uint32_t>();
        kv_type_[key] = type;
        auto scalar = [&](uint32_t t) {
            switch (t) {
            case T_U8: kv_u64_[key] = r.get<uint8_t>(); break;
            case T_I8: kv_u64_[key] = (uint64_t)(int64_t)r.get<int8_t>(); break;
            case T_U16: kv_u64_[key] = r.get<uint16_t>(); break;
            case T_I16: kv_u64_[key] = (uint64_t)(int64_t)r.get<int16_t>(); break;
            case T_U32: kv_u64_[key] = r.get<uint32_t>(); break;
            case T_I32: kv_u64_[key] = (uint64_t)(int64_t)r.get<int32_t>(); break;
            case T_BOOL: kv_u64_[key] = r.get<uint8_t>() != 0; break;
            case T_U64: kv_u64_[key] = r.get<uint64_t>(); break;
            case T_I64: kv_u64_[key] = (uint64_t)r.get<int64_t>(); break;
            case T_F32: kv_f64_[key] = r.get<float>(); break;
            case T_F64: kv_f64_[key] = r.get<double>(); break;
            default: r.ok = false;
            }
        };
        if (type == T_STR) {
            kv_str_[key] = r.str();
        } else if (type == T_ARR) {
            const uint32_t et = r.get<uint32_t>();
            const uint64_t n = r.get<uint64_t>();
            if (et == T_STR) {
                std::vector<std::string> &v = kv_strs_[key];
                if (n > (uint64_t)(r.end - r.p) / 8) { r.ok = false; break; }
                v.reserve((size_t)n);
                for (uint64_t k = 0; k < n && r.ok; k++) v.push_back(r.str());
            } else {
                const size_t es = scalar_size(et);
                if (!es || n > (uint64_t)(r.end - r.p) / es) { r.ok = false; break; }
                r.p += n * es; // numeric arrays (scores, token types) are not needed on this path
            }
        } else {
            scalar(type);
        }
    }
    tensors_.reserve((size_t)n_tensors);
    for (uint64_t i = 0; i < n_tensors && r.ok; i++) {
        gguf_tensor_info t;
        t.name = r.str();
        const uint32_t nd = r.get<uint32_t>();
        if (nd > 4) { r.ok = false; break; }
        uint64_t n = 1;
        for (uint32_t d = 0; d < nd; d++) {
            t.dims.push_back((int64_t)r.get<uint64_t>());
            n *= (uint64_t)t.dims.back();
        }
        t.type = (int32_t)r.get<uint32_t>();
        t.offset = r.get<uint64_t>();
        t.nbytes = type_bytes(t.type, n);
        tensors_.push_back(t);
    }
    if (!r.ok) { err_ = "truncated or malformed GGUF header"; return -3; }
    uint64_t align = 32;
    get_u64("general.alignment", &align);
    if (align == 0 || (align & (align - 1))) { err_ = "bad general.alignment"; return -3; }
    data_off_ = ((uint64_t)(r.p - blob_.data()) + align - 1) / align * align;
    for (const gguf_tensor_info &t : tensors_) {
        if (t.nbytes == 0) continue; // a type this path does not read: reported when it is asked for
        if (data_off_ + t.offset + t.nbytes > blob_.size()) { err_ = "tensor " + t.name + " runs past the end of the file"; return -3; }
    }
    return 0;
}

bool gguf_file::get_u64(const std::string &key, uint64_t *v) const {
    auto it = kv_u64_.find(key);
    if (it == kv_u64_.end()) return false;
    *v = it->second;
    return true;
}
bool gguf_file::get_f64(const std::string &key, double *v) const {
    auto it = kv_f64_.find(key);
    if (it != kv_f64_.end()) { *v = it->second; return true; }
    uint64_t u;
    if (get_u64(key, &u)) { *v = (double)u; return true; }
    return false;
}
bool gguf_file::get_str(const std::string &key, std::string *v) const {
    auto it = kv_str_.find(key);
    if (it == kv_str_.end()) return false;
    *v = it->second;
    return true;
}
const std::vector<std::string> *gguf_file::get_str_array(const std::string &key) const {
    auto it = kv_strs_.find(key);
    return it == kv_strs_.end() ? nullptr : &it->second;
}
const gguf_tensor_info *gguf_file::find(const std::string &name) const {
    for (const gguf_tensor_info &t : tensors_)
        if (t.name == name) return &t;
    return nullptr;
}

token_id gemma_tokenizer::find_token_id(const std::string &token) const {
    for (size_t i = 0; i < tokens.size(); i++)
        if (tokens[i] == token) return (token_id)i;
    return -1;
}
static void remove_wunderline(std::string &s) { // ref: src/gemma_model.cpp:753-760
    static const std::string underline = "\xE2\x96\x81"; // U+2581
    size_t pos = 0;
    while ((pos = s.find(underline, pos)) != std::string::npos) s.replace(pos, underline.length(), " ");
}
std::string gemma_tokenizer::token_text(token_id id) const {
    std::string t = find_token(id);
    remove_wunderline(t);
    return t;
}
std::string gemma_tokenizer::tokens_text(const std::vector<token_id> &ids) const {
    std::string t;
    for (token_id id : ids) t += find_token(id);
    const size_t b = t.find("<bos>");
    if (b != std::string::npos) t.replace(b, 5, "");
    remove_wunderline(t);
    return t;
}

int gguf_gemma_config(const gguf_file &f, int32_t n_ctx, gemma_config *cfg, std::string *err) {
    auto fail = [&](const std::string &m) { if (err) *err = m; return -1; };
    uint64_t L, E, H, Hkv, F = 0, ctx = 0, klen = 0;
    double eps = 1e-6;
    if (!f.get_u64("gemma.block_count", &L) || !f.get_u64("gemma.embedding_length", &E) ||
        !f.get_u64("gemma.attention.head_count", &H) || !f.get_u64("gemma.attention.head_count_kv", &Hkv))
        return fail("missing gemma.* hyper-parameters (block_count / embedding_length / attention.head_count[_kv])");
    f.get_f64("gemma.attention.layer_norm_rms_epsilon", &eps);
    f.get_u64("gemma.context_length", &ctx);
    f.get_u64("gemma.feed_forward_length", &F);
    const gguf_tensor_info *emb = f.find("token_embd.weight"), *gate = f.find("blk.0.ffn_gate.weight");
    if (!emb || emb->dims.size() != 2) return fail("token_embd.weight missing");
    if (!F && gate && gate->dims.size() == 2) F = (uint64_t)gate->dims[1];
    if (!f.get_u64("gemma.attention.key_length", &klen)) klen = H ? E / H : 0; // the reference's formula as the fallback
    if (!H || !Hkv || !klen || !F) return fail("bad hyper-parameters");
    memset(cfg, 0, sizeof(*cfg));
    cfg->n_embd = (int32_t)E; cfg->n_layer = (int32_t)L; cfg->n_head = (int32_t)H; cfg->n_head_kv = (int32_t)Hkv;
    cfg->head_dim = (int32_t)klen; cfg->n_ff = (int32_t)F; cfg->n_vocab = (int32_t)emb->dims[1];
    cfg->n_ctx = n_ctx > 0 ? n_ctx : (int32_t)(ctx ? ctx : 2048);
    cfg->eps = (float)eps; cfg->rope_base = 10000.0f; // ref: src/macro.h:12-18
    double base;
    if (f.get_f64("gemma.rope.freq_base", &base)) cfg->rope_base = (float)base;
    cfg->wtype = emb->type;
    // Q4_0 / Q8_0: the resident decode path; Q4_K / Q6_K (the token_embd of a q4_k_m file is Q6_K): K-quant mode
    if (cfg->wtype != GGB_TYPE_Q4_0 && cfg->wtype != GGB_TYPE_Q8_0 && cfg->wtype != GGB_TYPE_Q4_K && cfg->wtype != GGB_TYPE_Q6_K)
        return fail("token_embd.weight has ggml type " + std::to_string(emb->type) + ": this path runs Q4_0 / Q8_0 / Q4_K / Q6_K models");
    if ((cfg->wtype == GGB_TYPE_Q4_K || cfg->wtype == GGB_TYPE_Q6_K) && (E % 256 || F % 256 || (H * klen) % 256))
        return fail("K-quant model with a row length that is not a multiple of 256");
    if ((int64_t)E != emb->dims[0]) return fail("token_embd.weight width != gemma.embedding_length");
    return 0;
}

int gguf_gemma_tokenizer(const gguf_file &f, gemma_tokenizer *tok) {
    const std::vector<std::string> *t = f.get_str_array("tokenizer.ggml.tokens");
    if (!t) return -1;
    tok->tokens = *t;
    uint64_t v;
    if (f.get_u64("tokenizer.ggml.bos_token_id", &v)) tok->special_bos_id = (token_id)v;
    if (f.get_u64("tokenizer.ggml.eos_token_id", &v)) tok->special_eos_id = (token_id)v;
    return 0;
}

int gguf_bind_tensors(const gguf_file &f, gemma_model &m, std::string *err) {
    const gemma_config &c = m.config();
    auto bind = [&](const std::string &name, int id, bool matrix) -> int {
        const gguf_tensor_info *t = f.find(name);
        if (!t) { if (err) *err = "tensor " + name + " not in the file"; return -1; }
        // a Q4_0 / Q8_0 file holds ONE matrix type (the resident decode path streams it); a K-quant file (q4_k_m: Q4_K with
        // some Q6_K tensors) may mix Q4_K and Q6_K tensor by tensor
        const bool kq_file = c.wtype == GGB_TYPE_Q4_K || c.wtype == GGB_TYPE_Q6_K;
        const bool ok = !matrix ? t->type == 0 : kq_file ? (t->type == GGB_TYPE_Q4_K || t->type == GGB_TYPE_Q6_K) : t->type == (uint32_t)c.wtype;
        if (!ok) {
            if (err) *err = "tensor " + name + " has ggml type " + std::to_string(t->type) + " (expected " + std::to_string(matrix ? c.wtype : 0) + ")";
            return -1;
        }
        const int rc = m.set_tensor_typed(id, (int)t->type, f.data(*t), (int64_t)t->nbytes);
        if (rc && err) *err = "tensor " + name + ": shape does not match the hyper-parameters (" + std::to_string(t->nbytes) + " bytes)";
        return rc;
    };
    // ref: composite_model - token_embd.weight is also the output projection (:163)
    if (int rc = bind("token_embd.weight", GM_T_TOKEN_EMBD, true)) return rc;
    if (int rc = bind("output_norm.weight", GM_T_OUTPUT_NORM, false)) return rc;
    static const struct { const char *suffix; int id; bool matrix; } per_layer[] = {
        {"attn_norm.weight", GM_L_ATTN_NORM, false}, {"attn_q.weight", GM_L_ATTN_Q, true}, {"attn_k.weight", GM_L_ATTN_K, true},
        {"attn_v.weight", GM_L_ATTN_V, true}, {"attn_output.weight", GM_L_ATTN_OUT, true}, {"ffn_norm.weight", GM_L_FFN_NORM, false},
        {"ffn_gate.weight", GM_L_FFN_GATE, true}, {"ffn_up.weight", GM_L_FFN_UP, true}, {"ffn_down.weight", GM_L_FFN_DOWN, true}};
    for (int il = 0; il < c.n_layer; il++)
        for (const auto &pl : per_layer)
            if (int rc = bind("blk." + std::to_string(il) + "." + pl.suffix, GM_T_LAYER0 + GM_T_PER_LAYER * il + pl.id, pl.matrix)) return rc;
    return 0;
}
