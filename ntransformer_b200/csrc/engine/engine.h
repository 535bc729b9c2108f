// engine.h — text generation over the resident model: the reference's nt::Engine plain path
// (src/inference/engine.h:31-65, engine.cpp:16-23 load, :40-145 generate, :557-593 chat/benchmark, :595-607 stats).
// Speculative / self-speculative variants exist in the reference only to hide PCIe streaming and are out of scope.
#pragma once
#include <functional>
#include <string>
#include "model.h"
#include "text.h"
#include "tp_comm.h"

namespace nt { namespace b200 {

struct GenerateConfig {             // defaults of engine.h:17-26
    int max_tokens = 256;
    float temperature = 0.7f;
    int top_k = 40;
    float top_p = 0.9f;
    float repeat_penalty = 1.1f;
    int repeat_window = 64;
    uint64_t seed = 42;
    bool verbose = true;
    bool gpu_sampler = true;           // sample on the GPU when the settings allow it (csrc/sample.cu); --host-sampler / NT_B200_GPU_SAMPLER=0 turn it off
};
using TokenCallback = std::function<bool(const std::string& token, int token_id)>;

class Engine {
public:
    bool load(const std::string& model_path, int max_context = 4096);
    // Tensor-parallel rank of a one-process-per-GPU group (the CLI's --tp N forks the ranks): nccl_id is the 128-byte id made by
    // rank 0 (TPComm::unique_id).  Every rank runs the same generate() with the same seed; logits are bit-identical on all ranks,
    // so the replicas sample the same tokens without exchanging them.  Only rank 0 prints.
    bool load_tp(const std::string& model_path, int max_context, int tp_rank, int tp_size, const void* nccl_id);
    int tp_rank() const { return tp_rank_; }
    std::string generate(const std::string& prompt, const GenerateConfig& cfg, TokenCallback cb = nullptr);
    void chat(const GenerateConfig& cfg);
    void benchmark(const std::string& prompt, int n_tokens);
    const ModelConfig& config() const { return model_.config(); }
    Model& model() { return model_; }
    const Tokenizer& tokenizer() const { return tok_; }
    struct Stats {
        int prompt_tokens = 0, gen_tokens = 0;
        float prefill_ms = 0, decode_ms = 0;
    };
    const Stats& last_stats() const { return stats_; }
private:
    void print_stats(const Stats& s) const;
    std::unique_ptr<TPComm> comm_;       // declared before the model: destroyed after it (its graphs reference the communicator)
    Model model_;
    Tokenizer tok_;
    Stats stats_;
    int tp_rank_ = 0;
};

}}  // namespace nt::b200
