#include "engine.h"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>

namespace nt { namespace b200 {
using Clock = std::chrono::steady_clock;
static float ms_since(Clock::time_point t0) { return std::chrono::duration<float, std::milli>(Clock::now() - t0).count(); }

bool Engine::load(const std::string& path, int max_context) {
    if (!model_.load_gguf(path, max_context)) return false;
    tok_.init(model_.vocab(), model_.config().bos_token_id, model_.config().eos_token_id);
    return true;
}

bool Engine::load_tp(const std::string& path, int max_context, int tp_rank, int tp_size, const void* nccl_id) {
    tp_rank_ = tp_rank;
    if (tp_size <= 1) return load(path, max_context);
    if (!model_.load_gguf(path, max_context, tp_rank, tp_size)) return false;
    auto c = std::make_unique<TPComm>();
    if (!c->init(nccl_id, tp_rank, tp_size)) return false;
    {   // NCCL builds its channels lazily on the first collective: do that here, not inside the decode step's graph capture
        cudaStream_t st = model_.stream();
        float* tmp = nullptr;
        NT_CUDA_CHECK(cudaMalloc(&tmp, sizeof(float) * 1024 * (size_t)(tp_size + 1)));
        NT_CUDA_CHECK(cudaMemsetAsync(tmp, 0, sizeof(float) * 1024 * (size_t)(tp_size + 1), st));
        c->all_reduce_sum(tmp, 1024, st);
        c->all_gather(tmp, tmp + 1024, 1024, st);
        NT_CUDA_CHECK(cudaStreamSynchronize(st));
        NT_CUDA_CHECK(cudaFree(tmp));
    }
    model_.set_comm(c.get());
    comm_ = std::move(c);
    tok_.init(model_.vocab(), model_.config().bos_token_id, model_.config().eos_token_id);
    return true;
}

std::string Engine::generate(const std::string& prompt, const GenerateConfig& cfg, TokenCallback cb) {
    Stats st;
    Sampler sampler;
    SamplerConfig sc;
    sc.temperature = cfg.temperature; sc.top_k = cfg.top_k; sc.top_p = cfg.top_p;
    sc.repeat_penalty = cfg.repeat_penalty; sc.repeat_window = cfg.repeat_window; sc.seed = cfg.seed;
    sampler.init(sc);

    std::vector<int> tokens = tok_.encode(prompt, true);
    st.prompt_tokens = (int)tokens.size();
    if (cfg.verbose) fprintf(stderr, "Prompt tokens: %d\n", st.prompt_tokens);
    if (st.prompt_tokens > model_.config().max_seq_len) {
        // the reference writes past its KV cache here (transformer.cpp:629-640 has no check); an error return instead of an abort
        fprintf(stderr, "Error: the prompt has %d tokens, the context holds %d (use --ctx-size)\n", st.prompt_tokens, model_.config().max_seq_len);
        stats_ = st;
        return std::string();
    }
    const int vocab = model_.config().vocab_size;
    std::vector<float> logits((size_t)vocab);
    // Greedy without a repeat penalty needs no logits on the host: argmax runs on the GPU (4 B D2H instead of 513 KB).
    const bool gpu_greedy = cfg.temperature <= 0.0f && cfg.repeat_penalty <= 1.0f;
    // Default: penalty + top-k/top-p sampling on the GPU with the host's mt19937 stream (4 B D2H per token, csrc/sample.cu);
    // settings the kernel does not cover, --host-sampler and NT_B200_GPU_SAMPLER=0 take the host path (sampler.cpp:47-117).
    const char* gs_env = getenv("NT_B200_GPU_SAMPLER");
    const bool gs_off = gs_env && std::string(gs_env) == "0";
    const bool gpu_sample = cfg.gpu_sampler && !gs_off && sample_topk_supported(vocab, cfg.temperature, cfg.top_k);
    auto next_from = [&](float* dev_logits) {
        if (gpu_greedy) return model_.argmax_last();
        if (gpu_sample) {
            const int window = cfg.repeat_penalty > 1.0f ? std::min((int)tokens.size(), cfg.repeat_window) : 0;
            const int id = model_.sample_last(cfg.temperature, cfg.top_k, cfg.top_p, cfg.repeat_penalty,
                                              tokens.data() + (tokens.size() - (size_t)window), window, sampler.draw());
            if (id >= 0) return id;
        }
        NT_CUDA_CHECK(cudaMemcpy(logits.data(), dev_logits, sizeof(float) * (size_t)vocab, cudaMemcpyDeviceToHost));
        sampler.apply_repeat_penalty(logits.data(), vocab, tokens);
        return sampler.sample(logits.data(), vocab);
    };

    model_.clear_kv();
    auto t0 = Clock::now();
    float* dl = model_.forward(tokens.data(), (int)tokens.size(), 0);
    st.prefill_ms = ms_since(t0);
    int next = next_from(dl);
    tokens.push_back(next);
    std::string out, piece = tok_.decode_token(next);
    out += piece;
    bool stop = false;
    if (cb) stop = !cb(piece, next);
    else if (cfg.verbose) { fputs(piece.c_str(), stdout); fflush(stdout); }

    if (!stop) {
        auto d0 = Clock::now();
        int pos = st.prompt_tokens;
        for (int i = 1; i < cfg.max_tokens; i++) {
            if (next == tok_.eos_id()) break;
            if (pos >= model_.config().max_seq_len) break;           // the reference silently overruns its cache here
            dl = model_.forward(&next, 1, pos++);
            next = next_from(dl);
            tokens.push_back(next);
            piece = tok_.decode_token(next);
            out += piece;
            st.gen_tokens++;
            if (cb) { if (!cb(piece, next)) break; }
            else if (cfg.verbose) { fputs(piece.c_str(), stdout); fflush(stdout); }
        }
        st.decode_ms = ms_since(d0);
    }
    stats_ = st;
    if (cfg.verbose) { fputc('\n', stdout); print_stats(st); }
    return out;
}

void Engine::chat(const GenerateConfig& cfg) {
    fprintf(stderr, "\n=== NTransformer Chat ===\nType your message and press Enter. Type 'quit' to exit.\n\n");
    std::string line;
    for (;;) {
        fputs("> ", stdout); fflush(stdout);
        if (!std::getline(std::cin, line)) break;
        if (line == "quit" || line == "exit") break;
        if (line.empty()) continue;
        generate(line, cfg);                                         // stateless turns, like the reference
        fputc('\n', stdout);
    }
}

void Engine::benchmark(const std::string& prompt, int n_tokens) {
    GenerateConfig cfg;
    cfg.max_tokens = n_tokens; cfg.temperature = 0.0f; cfg.verbose = false;
    fprintf(stderr, "=== Benchmark ===\nPrompt: \"%s\"\nMax tokens: %d\n", prompt.c_str(), n_tokens);
    auto t0 = Clock::now();
    std::string out = generate(prompt, cfg);
    fprintf(stderr, "Total time: %.1f ms\nOutput length: %zu chars\n", ms_since(t0), out.size());
    print_stats(stats_);                                             // (the reference prints no tok/s here, quirk Q5)
}

void Engine::print_stats(const Stats& s) const {
    fprintf(stderr, "\n--- Stats ---\n");
    fprintf(stderr, "Prompt: %d tokens, %.1f ms (%.1f tok/s)\n", s.prompt_tokens, s.prefill_ms, s.prompt_tokens / (s.prefill_ms / 1000.0f));
    fprintf(stderr, "Decode: %d tokens, %.1f ms (%.1f tok/s)\n", s.gen_tokens, s.decode_ms, s.gen_tokens / (s.decode_ms / 1000.0f));
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    fprintf(stderr, "VRAM: %.1f / %.1f GB\n", (tot - fr) / (1024.0 * 1024 * 1024), tot / (1024.0 * 1024 * 1024));
}

}}  // namespace nt::b200
