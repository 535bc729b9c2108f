// gguf.h — GGUF v2/v3 reader (mmap) for the resident engine.
// Behaviour mirrors the reference loader (src/model/loader.cpp:23-276, src/model/config.cpp:18-50):
// scalars are narrowed to int/float, the vocab size is overridden by the token array length, arrays other
// than tokens/scores/token_type/merges are skipped (the reference also skips `tokenizer.ggml.merges`; we keep it for the opt-in BPE path), tensor data starts at
// the header end rounded up to `general.alignment` (default 32).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>
#include "../nt_types.h"

namespace nt { namespace b200 {

struct ModelConfig {                       // reference src/model/config.h:16-58
    std::string architecture = "llama";
    std::string model_name = "unknown";
    int vocab_size = 32000, hidden_size = 4096, intermediate_size = 11008, n_layers = 32;
    int n_heads = 32, n_kv_heads = 32, head_dim = 128;
    float norm_eps = 1e-5f, rope_theta = 10000.0f, rope_freq_scale = 1.0f;
    bool rope_interleaved = false;         // never set from GGUF by the reference
    int max_seq_len = 4096;
    int bos_token_id = 1, eos_token_id = 2;
    void print() const;
};

struct GGUFTensorInfo {
    std::string name;
    std::vector<int64_t> shape;            // ne[0] (fastest) first, as stored
    uint32_t ggml_type = 0;
    DType dtype = DType::F32;
    uint64_t offset = 0;                   // relative to the data section
    size_t nbytes = 0;
};

struct GGUFVocab {
    std::vector<std::string> tokens;
    std::vector<float> scores;
    std::vector<int> token_types;
    std::vector<std::string> merges;      // "left right" pairs in rank order (tokenizer.ggml.merges); the reference skips this array
};

class GGUFFile {
public:
    using Value = std::variant<int, float, std::string, bool>;
    GGUFFile() = default;
    ~GGUFFile();
    GGUFFile(const GGUFFile&) = delete;
    GGUFFile& operator=(const GGUFFile&) = delete;

    bool open(const std::string& path);    // false (+ message on stderr) on any parse failure
    const ModelConfig& config() const { return config_; }
    const GGUFVocab& vocab() const { return vocab_; }
    const std::vector<GGUFTensorInfo>& tensors() const { return tensors_; }
    const GGUFTensorInfo* find(const std::string& name) const;
    const void* data(const GGUFTensorInfo& t) const;   // host pointer into the mapping; aborts if out of file
    const std::unordered_map<std::string, Value>& metadata() const { return meta_; }
    size_t file_size() const { return size_; }
    size_t data_offset() const { return data_offset_; }
    int fd() const { return fd_; }          // for pread: tensor t lives at file offset data_offset() + t.offset
    void print_info() const;

private:
    bool parse();
    void* map_ = nullptr;
    size_t size_ = 0, data_offset_ = 0;
    int fd_ = -1;
    std::string path_;
    ModelConfig config_;
    GGUFVocab vocab_;
    std::vector<GGUFTensorInfo> tensors_;
    std::unordered_map<std::string, size_t> index_;
    std::unordered_map<std::string, Value> meta_;
};

}}  // namespace nt::b200
