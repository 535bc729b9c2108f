// gguf.cpp — see gguf.h. Bounds-checked cursor over the mmap (the reference trusts the file): every count, shape and
// offset read from the file is checked against the bytes that remain before anything is allocated or dereferenced.
#include "gguf.h"

#include <cinttypes>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace nt { namespace b200 {

namespace {
struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    template <typename T> T get() {
        T v{};
        if (!ok || (size_t)(end - p) < sizeof(T)) { ok = false; return v; }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        uint64_t n = get<uint64_t>();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return {}; }
        std::string s(reinterpret_cast<const char*>(p), (size_t)n);
        p += n;
        return s;
    }
    void skip(size_t n) { if (!ok || n > (size_t)(end - p)) ok = false; else p += n; }
    size_t left() const { return (size_t)(end - p); }
    // n elements of es bytes each, without overflow
    void skip_n(uint64_t n, size_t es) { if (!ok || es == 0 || n > left() / es) ok = false; else p += (size_t)n * es; }
};

enum : uint32_t { T_U8, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64 };

size_t scalar_size(uint32_t t) {
    switch (t) {
        case T_U8: case T_I8: case T_BOOL: return 1;
        case T_U16: case T_I16: return 2;
        case T_U32: case T_I32: case T_F32: return 4;
        case T_U64: case T_I64: case T_F64: return 8;
        default: return 0;
    }
}

// Scalars narrowed to int/float exactly like loader.cpp:197-215 (quirk Q8).
GGUFFile::Value read_scalar(Cursor& c, uint32_t t) {
    switch (t) {
        case T_U8: return (int)c.get<uint8_t>();
        case T_I8: return (int)c.get<int8_t>();
        case T_U16: return (int)c.get<uint16_t>();
        case T_I16: return (int)c.get<int16_t>();
        case T_U32: return (int)c.get<uint32_t>();
        case T_I32: return (int)c.get<int32_t>();
        case T_U64: return (int)c.get<uint64_t>();
        case T_I64: return (int)c.get<int64_t>();
        case T_F32: return c.get<float>();
        case T_F64: return (float)c.get<double>();
        case T_BOOL: return c.get<uint8_t>() != 0;
        case T_STR: return c.str();
        default: c.ok = false; return 0;
    }
}

void skip_value(Cursor& c, uint32_t t) {
    if (t == T_STR) { c.str(); return; }
    if (t == T_ARR) {
        uint32_t et = c.get<uint32_t>();
        uint64_t n = c.get<uint64_t>();
        if (size_t es = scalar_size(et)) { c.skip_n(n, es); return; }
        if (et == T_ARR) { c.ok = false; return; }           // nested arrays: not produced by any writer we load; bounds the recursion
        if (n > c.left()) { c.ok = false; return; }          // every element takes >= 1 byte
        for (uint64_t i = 0; i < n && c.ok; i++) skip_value(c, et);
        return;
    }
    size_t s = scalar_size(t);
    if (!s) c.ok = false; else c.skip(s);
}

template <typename T>
T meta_get(const std::unordered_map<std::string, GGUFFile::Value>& kv, const std::string& key, T dflt) {
    auto it = kv.find(key);
    if (it == kv.end()) return dflt;
    if (auto* v = std::get_if<T>(&it->second)) return *v;
    return dflt;
}
}  // namespace

void ModelConfig::print() const {          // same lines as reference config.cpp:52-64
    fprintf(stderr, "=== Model Config ===\n");
    fprintf(stderr, "Architecture: %s\n", architecture.c_str());
    fprintf(stderr, "Name: %s\n", model_name.c_str());
    fprintf(stderr, "Vocab: %d, Hidden: %d, Intermediate: %d\n", vocab_size, hidden_size, intermediate_size);
    fprintf(stderr, "Layers: %d, Heads: %d, KV Heads: %d, Head dim: %d\n", n_layers, n_heads, n_kv_heads, head_dim);
    fprintf(stderr, "Max seq: %d, Norm eps: %e\n", max_seq_len, norm_eps);
    fprintf(stderr, "RoPE theta: %.1f, GQA: %s (group=%d)\n", rope_theta, n_kv_heads < n_heads ? "yes" : "no",
            n_kv_heads ? n_heads / n_kv_heads : 0);
    fprintf(stderr, "BOS: %d, EOS: %d\n", bos_token_id, eos_token_id);
}

GGUFFile::~GGUFFile() {
    if (map_ && map_ != MAP_FAILED) munmap(map_, size_);
    if (fd_ >= 0) close(fd_);
}

bool GGUFFile::open(const std::string& path) {
    path_ = path;
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { fprintf(stderr, "Failed to open %s\n", path.c_str()); return false; }
    struct stat st;
    if (fstat(fd_, &st) != 0) { fprintf(stderr, "Failed to stat %s\n", path.c_str()); return false; }
    size_ = (size_t)st.st_size;
    map_ = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (map_ == MAP_FAILED) { fprintf(stderr, "Failed to mmap %s\n", path.c_str()); map_ = nullptr; return false; }
    madvise(map_, size_, MADV_SEQUENTIAL);
    if (!parse()) { fprintf(stderr, "Failed to parse GGUF header\n"); return false; }
    return true;
}

bool GGUFFile::parse() {
    const uint8_t* base = static_cast<const uint8_t*>(map_);
    Cursor c{base, base + size_};
    uint32_t magic = c.get<uint32_t>();
    if (!c.ok || magic != 0x46554747u) { fprintf(stderr, "Invalid GGUF magic: 0x%08X (expected 0x46554747)\n", magic); return false; }
    uint32_t version = c.get<uint32_t>();
    if (version < 2 || version > 3) { fprintf(stderr, "Unsupported GGUF version: %u\n", version); return false; }
    uint64_t n_tensors = c.get<uint64_t>(), n_kv = c.get<uint64_t>();
    if (!c.ok) return false;
    // a key needs >= 12 bytes (length + type), a tensor record >= 24 (name length, nd, type, offset): cap the counts by the file
    if (n_kv > c.left() / 12 || n_tensors > c.left() / 24) { fprintf(stderr, "GGUF: header counts exceed the file size\n"); return false; }
    fprintf(stderr, "GGUF v%u: %" PRIu64 " tensors, %" PRIu64 " metadata entries\n", version, n_tensors, n_kv);

    for (uint64_t i = 0; i < n_kv && c.ok; i++) {
        std::string key = c.str();
        uint32_t type = c.get<uint32_t>();
        if (type == T_ARR) {
            uint32_t et = c.get<uint32_t>();
            uint64_t n = c.get<uint64_t>();
            if (!c.ok) break;
            if (key == "tokenizer.ggml.tokens" && et == T_STR) {
                if (n > c.left() / 8) { c.ok = false; break; }          // a string takes >= 8 bytes
                vocab_.tokens.reserve((size_t)n);
                for (uint64_t j = 0; j < n && c.ok; j++) vocab_.tokens.push_back(c.str());
            } else if (key == "tokenizer.ggml.merges" && et == T_STR) {
                if (n > c.left() / 8) { c.ok = false; break; }
                vocab_.merges.reserve((size_t)n);
                for (uint64_t j = 0; j < n && c.ok; j++) vocab_.merges.push_back(c.str());
            } else if (key == "tokenizer.ggml.scores" && et == T_F32) {
                if (n > c.left() / 4) { c.ok = false; break; }
                vocab_.scores.reserve((size_t)n);
                for (uint64_t j = 0; j < n && c.ok; j++) vocab_.scores.push_back(c.get<float>());
            } else if (key == "tokenizer.ggml.token_type" && et == T_I32) {
                if (n > c.left() / 4) { c.ok = false; break; }
                vocab_.token_types.reserve((size_t)n);
                for (uint64_t j = 0; j < n && c.ok; j++) vocab_.token_types.push_back(c.get<int32_t>());
            } else if (size_t es = scalar_size(et)) {
                c.skip_n(n, es);
            } else {
                if (et == T_ARR || n > c.left()) { c.ok = false; break; }
                for (uint64_t j = 0; j < n && c.ok; j++) skip_value(c, et);
            }
        } else {
            meta_[key] = read_scalar(c, type);
        }
    }
    if (!c.ok) return false;

    // metadata -> config (config.cpp:18-50)
    config_.architecture = meta_get<std::string>(meta_, "general.architecture", "llama");
    config_.model_name = meta_get<std::string>(meta_, "general.name", "unknown");
    const std::string pfx = config_.architecture + ".";
    config_.vocab_size = meta_get<int>(meta_, pfx + "vocab_size", config_.vocab_size);
    config_.hidden_size = meta_get<int>(meta_, pfx + "embedding_length", config_.hidden_size);
    config_.intermediate_size = meta_get<int>(meta_, pfx + "feed_forward_length", config_.intermediate_size);
    config_.n_layers = meta_get<int>(meta_, pfx + "block_count", config_.n_layers);
    config_.n_heads = meta_get<int>(meta_, pfx + "attention.head_count", config_.n_heads);
    config_.n_kv_heads = meta_get<int>(meta_, pfx + "attention.head_count_kv", config_.n_heads);
    if (config_.n_heads <= 0) { fprintf(stderr, "GGUF: attention.head_count must be positive\n"); return false; }
    config_.head_dim = config_.hidden_size / config_.n_heads;
    config_.max_seq_len = meta_get<int>(meta_, pfx + "context_length", config_.max_seq_len);
    config_.norm_eps = meta_get<float>(meta_, pfx + "attention.layer_norm_rms_epsilon", config_.norm_eps);
    config_.rope_theta = meta_get<float>(meta_, pfx + "rope.freq_base", config_.rope_theta);
    config_.bos_token_id = meta_get<int>(meta_, "tokenizer.ggml.bos_token_id", config_.bos_token_id);
    config_.eos_token_id = meta_get<int>(meta_, "tokenizer.ggml.eos_token_id", config_.eos_token_id);
    if (!vocab_.tokens.empty() && (int)vocab_.tokens.size() != config_.vocab_size) config_.vocab_size = (int)vocab_.tokens.size();

    tensors_.resize((size_t)n_tensors);
    for (uint64_t i = 0; i < n_tensors && c.ok; i++) {
        GGUFTensorInfo& t = tensors_[(size_t)i];
        t.name = c.str();
        uint32_t nd = c.get<uint32_t>();
        if (nd < 1 || nd > 4) { c.ok = false; break; }        // GGML_MAX_DIMS = 4; a tensor has at least one dimension
        uint64_t n = 1;
        for (uint32_t d = 0; d < nd && c.ok; d++) {
            const uint64_t e = c.get<uint64_t>();
            // element counts stay far below 2^40 (a 70B embedding is 2^30): rejects zero, negative-as-int64 and overflowing shapes
            if (e == 0 || e > ((uint64_t)1 << 40) || n > ((uint64_t)1 << 40) / e) { c.ok = false; break; }
            t.shape.push_back((int64_t)e);
            n *= e;
        }
        if (!c.ok) break;
        t.ggml_type = c.get<uint32_t>();
        t.dtype = ggml_to_dtype(t.ggml_type);
        t.offset = c.get<uint64_t>();
        const size_t blk = dtype_block_size(t.dtype);
        if (blk == 0 || (uint64_t)t.shape[0] % blk != 0) {       // quant blocks run along dimension 0 (the row)
            fprintf(stderr, "GGUF: tensor '%s': row length %lld is not a multiple of the %zu-element block\n", t.name.c_str(), (long long)t.shape[0], blk);
            return false;
        }
        t.nbytes = dtype_row_size(t.dtype, (size_t)n);
        index_[t.name] = (size_t)i;
    }
    if (!c.ok) return false;
    int alignment = meta_get<int>(meta_, "general.alignment", 32);
    if (alignment <= 0 || (alignment & (alignment - 1))) alignment = 32;
    size_t header = (size_t)(c.p - base);
    data_offset_ = (header + (size_t)alignment - 1) & ~((size_t)alignment - 1);
    // every tensor must lie inside the mapping (overflow-safe): checked here once, so data() can never hand out a wild pointer
    if (data_offset_ > size_) {
        if (!tensors_.empty()) { fprintf(stderr, "GGUF: tensor data starts beyond the end of the file\n"); return false; }
        data_offset_ = size_;
    }
    const size_t avail = size_ - data_offset_;
    for (const GGUFTensorInfo& t : tensors_) {
        if (t.offset > avail || t.nbytes > avail - (size_t)t.offset) {
            fprintf(stderr, "GGUF: tensor '%s' extends beyond the file (offset %" PRIu64 ", %zu bytes, %zu available)\n", t.name.c_str(),
                    (uint64_t)t.offset, t.nbytes, avail);
            return false;
        }
    }
    return true;
}

const GGUFTensorInfo* GGUFFile::find(const std::string& name) const {
    auto it = index_.find(name);
    return it == index_.end() ? nullptr : &tensors_[it->second];
}

const void* GGUFFile::data(const GGUFTensorInfo& t) const {
    const size_t avail = size_ - data_offset_;
    if (t.offset > avail || t.nbytes > avail - (size_t)t.offset) {      // same fatal condition as loader.cpp:247-255; parse() already rejects such files
        fprintf(stderr, "\nERROR: Tensor '%s' extends beyond file! data_offset=%zu, tensor_offset=%zu, nbytes=%zu, file_size=%zu\n",
                t.name.c_str(), data_offset_, (size_t)t.offset, t.nbytes, size_);
        abort();
    }
    return static_cast<const uint8_t*>(map_) + data_offset_ + t.offset;
}

void GGUFFile::print_info() const {
    fprintf(stderr, "=== GGUF File: %s ===\n", path_.c_str());
    fprintf(stderr, "File size: %.2f GB\n", size_ / (1024.0 * 1024 * 1024));
    fprintf(stderr, "Tensor data: %.2f GB at offset 0x%zX\n", (size_ - data_offset_) / (1024.0 * 1024 * 1024), data_offset_);
    fprintf(stderr, "Tensors: %zu\n", tensors_.size());
    fprintf(stderr, "Vocab: %zu tokens\n", vocab_.tokens.size());
}

}}  // namespace nt::b200
