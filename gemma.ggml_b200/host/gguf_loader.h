// gguf_loader.h — GGUF file reader + Gemma tokenizer.  Mirrors what the reference does through
// ggml's gguf_* API (ref: src/gemma_model.cpp:19-51 load_model_from_file, :583-648 header KV / tensor table,
// :403-415 hyper-parameters, :159-194 tensor binding by name, load_tokenizer) with a self-contained parser: the file
// is read once, the 2 + 9 L tensors are handed to gemma_model::set_tensor (ggml AoS blocks, uploaded and repacked on
// the device), the vocabulary stays on the host for print_token (ref: :749-793).
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "gemma_model.h"

struct gguf_tensor_info {
    std::string name;
    std::vector<int64_t> dims; // ggml order: dims[0] = columns (K)
    int32_t type = 0;          // ggml_type
    uint64_t offset = 0;       // from the start of the data section
    uint64_t nbytes = 0;
};

class gguf_file {
  public:
    // returns 0 or a negative error; error() says why
    int open(const char *path);
    const std::string &error() const { return err_; }
    uint32_t version() const { return version_; }

    bool has(const std::string &key) const { return kv_type_.count(key) != 0; }
    // scalars (any integer / float / bool type is widened)
    bool get_u64(const std::string &key, uint64_t *v) const;
    bool get_f64(const std::string &key, double *v) const;
    bool get_str(const std::string &key, std::string *v) const;
    const std::vector<std::string> *get_str_array(const std::string &key) const;
    size_t n_kv() const { return kv_type_.size(); }

    const std::vector<gguf_tensor_info> &tensors() const { return tensors_; }
    const gguf_tensor_info *find(const std::string &name) const;
    const uint8_t *data(const gguf_tensor_info &t) const { return blob_.data() + data_off_ + t.offset; }

  private:
    std::string err_;
    uint32_t version_ = 0;
    std::map<std::string, uint32_t> kv_type_;
    std::map<std::string, uint64_t> kv_u64_;
    std::map<std::string, double> kv_f64_;
    std::map<std::string, std::string> kv_str_;
    std::map<std::string, std::vector<std::string>> kv_strs_;
    std::vector<gguf_tensor_info> tensors_;
    std::vector<uint8_t> blob_;
    uint64_t data_off_ = 0;
};

// ref: gemma_tokenizer (src/gemma_model.h, src/gemma_model.cpp:749-793)
struct gemma_tokenizer {
    std::vector<std::string> tokens;
    token_id special_bos_id = 2, special_eos_id = 1; // ref defaults (tokenizer.ggml.bos_token_id / eos_token_id)
    std::string find_token(token_id id) const { return id >= 0 && (size_t)id < tokens.size() ? tokens[(size_t)id] : std::string(); }
    token_id find_token_id(const std::string &token) const;
    // the text of a token / a token list: "▁" -> " " and a leading "<bos>" removed (ref: print_token / print_tokens)
    std::string token_text(token_id id) const;
    std::string tokens_text(const std::vector<token_id> &ids) const;
};

// ref: init_hyper_param (:403-415).  head_dim comes from gemma.attention.key_length when present (the reference's
// n_embd / n_head is wrong for Gemma-7B); n_ctx: 0 = the file's gemma.context_length
int gguf_gemma_config(const gguf_file &f, int32_t n_ctx, gemma_config *cfg, std::string *err);
// ref: load_tokenizer
int gguf_gemma_tokenizer(const gguf_file &f, gemma_tokenizer *tok);
// ref: composite_model (:159-194): the tensors by the reference's names -> gemma_model::set_tensor
int gguf_bind_tensors(const gguf_file &f, gemma_model &m, std::string *err);
