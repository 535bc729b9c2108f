// text.h — tokenizer and sampler of the CLI path.
// Behavioural mirror of the reference's approximate BPE (src/inference/tokenizer.cpp:64-314: GPT-2 byte map
// auto-detected by the presence of the encoded space, greedy longest-match then highest-score pair merges,
// control/unused tokens decode to "") and of its CPU sampler (src/inference/sampler.cpp:18-117:
// repeat penalty over the last `repeat_window` tokens, temperature, top-k, top-p, std::mt19937).
#pragma once
#include <cstdint>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>
#include "gguf.h"

namespace nt { namespace b200 {

class Tokenizer {
public:
    void init(const GGUFVocab& vocab, int bos_id, int eos_id);
    std::vector<int> encode(const std::string& text, bool add_bos = true) const;
    std::string decode(const std::vector<int>& ids) const;
    std::string decode_token(int id) const;
    int bos_id() const { return bos_; }
    int eos_id() const { return eos_; }
    int vocab_size() const { return (int)tokens_.size(); }
    bool gpt2() const { return gpt2_; }
    // Opt-in (CLI --bpe-merges, NT_B200_BPE_MERGES=1): rank-ordered byte-level BPE over tokenizer.ggml.merges with a Llama-3 style
    // pre-tokeniser and literal special tokens — what llama.cpp does for this vocabulary family.  Off by default because the
    // reference ignores the merges (loader.cpp skips the array, tokenizer.cpp:101-217 merges by score) and parity is against it.
    void set_use_merges(bool on) { use_merges_ = on; }
    bool merges_active() const { return use_merges_ && gpt2_ && !merge_rank_.empty(); }
private:
    int byte_token(uint8_t b) const;
    void encode_merges(const std::string& text, std::vector<int>& out) const;
    void bpe_word(const std::string& word, std::vector<int>& out) const;
    std::unordered_map<std::string, int> merge_rank_;     // "left\x01right" -> rank
    std::vector<std::pair<std::string, int>> specials_;   // control tokens that may appear literally in text, longest first
    bool use_merges_ = false;
    std::vector<std::string> tokens_;
    std::vector<float> scores_;
    std::vector<int> types_;
    std::unordered_map<std::string, int> ids_;
    int bos_ = 1, eos_ = 2;
    bool gpt2_ = false;
};

struct SamplerConfig {
    float temperature = 0.7f;
    int top_k = 40;
    float top_p = 0.9f;
    float repeat_penalty = 1.1f;
    int repeat_window = 64;
    uint64_t seed = 42;
};

class Sampler {
public:
    void init(const SamplerConfig& c) { cfg_ = c; rng_.seed(c.seed); }
    static int argmax(const float* logits, int n);
    void apply_repeat_penalty(float* logits, int n, const std::vector<int>& recent) const;
    int sample(const float* logits, int n);
    // The uniform variate Sampler::sample would draw next (sampler.cpp:105-106): lets the GPU sampler consume the same stream.
    float draw() { std::uniform_real_distribution<float> dist(0.0f, 1.0f); return dist(rng_); }
    const SamplerConfig& config() const { return cfg_; }
private:
    SamplerConfig cfg_;
    std::mt19937 rng_;
    std::vector<std::pair<float, int>> cand_;
};

}}  // namespace nt::b200
