// model.h — resident Llama-family forward on one GPU (or one tensor-parallel rank).
//
// Host-side counterpart of the reference's nt::Transformer resident branch
// (src/model/transformer.cpp:59-126 load, :286-391 layers/buffers, :604-669 forward) with the per-layer
// launch sequence of Attention::forward (src/model/attention.cpp:120-211) and FFN::forward
// (src/model/ffn.cpp:85-134) re-expressed as fused sm_100a launches captured in a CUDA graph:
//   reference: 15 launches / layer / token, CPU embedding dequant, 2 sync H2D + 1 sync D2H per token;
//   here:      10 launches / layer / token in one graph replay, embedding gathered on the GPU.
// The streaming / tiered / speculative branches of the reference are out of scope.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <memory>
#include "../kernels_internal.h"
#include "gguf.h"
#include "decode_mega.h"
#include "peer_xchg.h"

namespace nt { namespace b200 {

class TPComm;    // NCCL communicator (engine/tp_comm.cpp); null when tp_size == 1

struct Weight {
    const void* ptr = nullptr;   // device, GGUF block layout, this rank's shard
    DType dtype = DType::F32;
    int rows = 0, cols = 0;      // shard shape (out_features, in_features)
    size_t pitch = 0;            // bytes between rows
    bool owned = false;
    size_t bytes() const { return (size_t)rows * dtype_row_size(dtype, (size_t)cols); }   // algorithmic bytes
};

struct LayerWeights {
    const float* attn_norm = nullptr;
    const float* ffn_norm = nullptr;
    Weight wq, wk, wv, wo, gate, up, down;
};

class Model {
public:
    Model() = default;
    ~Model();
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    // (a) from a GGUF file: parse, shard for (tp_rank, tp_size), upload, allocate, capture.
    bool load_gguf(const std::string& path, int max_context, int tp_rank = 0, int tp_size = 1);
    // (b) from caller-owned device tensors (synthetic benchmarks): init -> set_tensor* -> finalize.
    void init(const ModelConfig& cfg, int tp_rank = 0, int tp_size = 1);
    bool set_tensor(const std::string& gguf_name, const void* dev_ptr, DType dtype, size_t row_pitch = 0);
    bool finalize();

    void set_comm(TPComm* comm) { comm_ = comm; }

    // Runs seq_len tokens at positions start_pos.. and returns DEVICE logits [vocab] of the last token
    // (same contract as Transformer::forward, transformer.cpp:604-669; synchronises the stream).
    float* forward(const int* tokens, int seq_len, int start_pos);
    // Asynchronous variant used by the benchmark: no host sync.
    void forward_async(const int* tokens, int seq_len, int start_pos);
    // Greedy next token computed on the GPU from the last logits (argmax, lowest index on ties like Sampler::argmax).
    int argmax_last();
    int sync();                            // waits for the model's stream: returns the cudaError_t; reports a persistent-kernel time-out
    // Samples the next token from the last logits on the GPU (csrc/sample.cu; penalty is applied in place to the device
    // logits).  recent_window: the ids the repeat penalty looks at, oldest first.  Returns -1 when the settings are not
    // covered (top_k <= 0, > 1024 or >= vocab, temperature <= 0): the caller samples on the host.
    int sample_last(float temperature, int top_k, float top_p, float repeat_penalty, const int* recent_window, int n_window, float r);

    const ModelConfig& config() const { return cfg_; }
    const GGUFVocab& vocab() const { return vocab_; }
    cudaStream_t stream() const { return stream_; }
    float* logits_device() const { return logits_; }
    int tp_rank() const { return tp_rank_; }
    int tp_size() const { return tp_size_; }
    // Algorithmic bytes this rank reads per decoded token at context length ctx (SURVEY §8d B_tok).
    size_t bytes_per_token(int ctx) const;
    size_t weight_bytes() const;
    double load_seconds() const { return load_seconds_; }     // wall time of load_gguf (parse + plan + pipelined upload + buffers)
    void set_use_graph(bool on) { use_graph_ = on; }
    void set_use_pdl(bool on) { use_pdl_ = on; }
    // Prompts of at least n tokens go through the batched tensor-core prefill when the weights allow it (0 = never).
    void set_prefill_min_tokens(int n) { prefill_min_tokens_ = n; }
    bool batched_prefill_ok(int seq_len, int start_pos) const;
    void clear_kv();
    // Opt-in (also NT_B200_MEGAKERNEL=1): run the decode step as one persistent kernel (engine/decode_mega.h) instead of
    // the graph of fused launches.  Falls back to the graph path, with a note on stderr, when a shape is not covered.
    void set_use_megakernel(bool on) { use_mega_ = on; }
    int tp_exchange_kind() const { return tp_size_ == 1 || !xchg_tried_ ? 0 : (xchg_ ? 2 : 1); }
    bool megakernel_active();                 // true once the persistent kernel has been built for this model
    // Debug read-back of the persistent kernel's working buffers ("hid0", "hid1", "q", "attn", "act", "slots"): device pointer.
    const float* mega_debug_buffer(const char* name, size_t* count);
    void mega_trace(bool on);                  // record the persistent kernel's phase timeline (debug/tuning)
    size_t mega_trace_read(unsigned long long* out_host, size_t cap);   // [4 CTAs][phases][start, work done, barrier passed] ns
    int mega_plan_kinds(int* kinds, int cap);  // phase kinds of the per-token program; returns their number (0 when inactive)

private:
    bool ensure_mega();
    void run_step_mega(bool with_head);
    void step_body(cudaStream_t s);      // embedding + all layers for the token/position in step_dev_
    void step_head(cudaStream_t s, bool from_chain = false);   // final norm + LM head (+ all-gather under TP); from_chain: norm + quantiser in the GEMV prologue
    bool chain_ok() const;               // every projection on the TMA/dp4a GEMV: the fused launch chain applies
    void run_step(bool with_head);
    void matvec(const Weight* const* ws, float* const* ys, int n, const float* x, const float* norm_w, GemvEpilogue ep,
                cudaStream_t s);
    void reduce_residual(float* partial, cudaStream_t s);
    // All prompt tokens at once: tcgen05 GEMMs (csrc/prefill_gemm.cu) + batched norm/rope/KV-write/causal attention,
    // instead of the reference's per-token loop (transformer.cpp:604-669 runs every layer's GEMVs once per token).
    void prefill_batched(const int* tokens, int seq_len, int start_pos);
    void ensure_prefill_buffers(int tokens);
    bool f16_direct(const Weight& w) const;   // F16 rows usable by TMA in place
    // C[T, w.rows] (+)= split(A)[T, w.cols] . W^T for a weight of any GGUF dtype (quantised: dequantise to hi/lo, two GEMMs)
    void prefill_gemm(float* C, const void* ws, const Weight& w, int T, bool add, cudaStream_t s);
    bool o_xq_fusable(const Weight& wo) const;
    void release_graphs();

    ModelConfig cfg_;
    GGUFVocab vocab_;
    int tp_rank_ = 0, tp_size_ = 1;
    int nh_l_ = 0, nkv_l_ = 0;           // local heads
    int inter_l_ = 0, vocab_l_ = 0;      // local FFN columns / vocab rows
    TPComm* comm_ = nullptr;

    Weight embd_, head_;
    const float* out_norm_ = nullptr;
    std::vector<LayerWeights> layers_;
    std::vector<void*> owned_;

    // device buffers
    cudaStream_t stream_ = nullptr;
    float *hidden_ = nullptr, *xnorm_ = nullptr, *q_ = nullptr, *k_ = nullptr, *v_ = nullptr, *attn_ = nullptr;
    float *act_ = nullptr, *up_ = nullptr, *part_ = nullptr, *logits_ = nullptr, *logits_l_ = nullptr, *attn_scratch_ = nullptr;
    void *xq_h_ = nullptr, *xq_a_ = nullptr, *xq_i_ = nullptr;
    unsigned* attn_tickets_ = nullptr;   // last-arriver tickets of the one-launch attention (zero between launches)
    int fuse_mask_ = 1;                  // bit 0: norm + quantiser in the GEMV prologues (default), bit 1: one-launch attention (opt-in: measured slower, profiles/r02_*), bit 2: RoPE + KV write inside the decode kernel, merge separate (NT_B200_FUSE)
    void *kc_ = nullptr, *vc_ = nullptr;
    int* step_dev_ = nullptr;            // [0] token, [1] position
    int* argmax_dev_ = nullptr;
    int* argmax_host_ = nullptr;
    int* recent_dev_ = nullptr;          // repeat-penalty window for the GPU sampler
    int recent_cap_ = 0;

    struct PrefillBuffers {
        int cap = 0;                     // tokens per chunk the buffers hold
        float *x = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *attn = nullptr;
        void *ws = nullptr, *ws2 = nullptr;   // F16 hi/lo split of the current GEMM input (ws2: SwiGLU output -> down projection)
        float *g = nullptr, *u = nullptr;     // gate / up activations (only when those weights are not F16: unfused SwiGLU)
        void *whi = nullptr, *wlo = nullptr;  // dequantised weight matrix as an F16 hi/lo pair (quantised models)
        int *tok = nullptr, *pos = nullptr;
    } pf_;
    // Prompts shorter than this replay token by token through the decode chain (exact-integer GEMVs); longer ones take the batched
    // tensor-core prefill, whose F32 accumulation inside the tensor cores truncates once per 16-deep MMA step: ~1e-5..1e-4 per GEMM,
    // ~3e-4 per layer on the random synthetic models (profiles/r02_parity_8b.txt) against ~1e-5 per layer for the GEMV path.
    int prefill_min_tokens_ = 32;

    bool use_graph_ = true;
    bool use_pdl_ = true;
    cudaGraphExec_t g_full_ = nullptr, g_body_ = nullptr;
    int n_full_ = 0, n_body_ = 0;        // kernels per graph replay
    bool finalized_ = false;
    double load_seconds_ = 0.0;

    // Tensor parallel: the o-projection / down-projection partials are summed through NVLink peer memory (engine/peer_xchg.h);
    // NT_B200_TP_NCCL=1 (or a failed peer mapping) keeps ncclAllReduce + add.
    std::unique_ptr<PeerXchg> xchg_;
    bool xchg_tried_ = false;
    void ensure_xchg();

    bool use_mega_ = false;
    bool mega_tried_ = false;
    std::unique_ptr<DecodeMega> mega_;
};

}}  // namespace nt::b200
