/*
 * nt_b200_engine.h — token-level C-ABI of the resident model (additions to include/ntransformer.h).
 *
 * Mirrors the reference's C++ surface nt::Transformer::{load,forward,config} (src/model/transformer.h:31-61,
 * forward = src/model/transformer.cpp:604-669): tokens in (HOST int32), logits out (HOST or DEVICE float32
 * [vocab]).  Weights come either from a GGUF file or from caller-owned device tensors in GGUF block layout
 * (synthetic benchmarks); tensor names are the reference's GGUF names (transformer.cpp:89-104, 286-322).
 * Tensor parallelism (new; the reference is single GPU): one process per GPU, ranks exchange a 128-byte
 * NCCL id through their own plumbing and call nt_tp_init before the first forward.
 */
#ifndef NT_B200_ENGINE_H
#define NT_B200_ENGINE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int vocab_size, hidden_size, intermediate_size, n_layers, n_heads, n_kv_heads, head_dim, max_seq_len;
    float norm_eps, rope_theta;
    int bos_token_id, eos_token_id;
} nt_model_config;

typedef void* nt_model_t;

/* GGUF path: returns NULL on failure.  max_context caps llama.context_length like the CLI's -c. */
nt_model_t nt_model_load_gguf(const char* path, int max_context, int tp_rank, int tp_size);
/* Device-pointer path: create -> set_tensor (every tensor, shapes already sharded for the rank) -> finalize */
nt_model_t nt_model_create(const nt_model_config* cfg, int tp_rank, int tp_size);
int  nt_model_set_tensor(nt_model_t m, const char* gguf_name, const void* dev_ptr, int dtype, size_t row_pitch);
int  nt_model_finalize(nt_model_t m);
void nt_model_free(nt_model_t m);
int  nt_model_get_config(nt_model_t m, nt_model_config* out);

/* forward(tokens[seq_len], start_pos): logits of the last token. logits_host may be NULL.  Returns 0. */
int  nt_model_forward(nt_model_t m, const int* tokens_host, int seq_len, int start_pos, float* logits_host);
/* asynchronous launch only (no sync, no copy); nt_model_sync waits for the model's stream */
int  nt_model_forward_async(nt_model_t m, const int* tokens_host, int seq_len, int start_pos);
int  nt_model_sync(nt_model_t m);
float* nt_model_logits_device(nt_model_t m);
void*  nt_model_stream(nt_model_t m);
int  nt_model_argmax(nt_model_t m);                 /* greedy token of the last logits, computed on the GPU */
void nt_model_clear_kv(nt_model_t m);
void nt_model_use_graph(nt_model_t m, int on);
/* prompts of >= n tokens use the batched tcgen05 prefill when every layer matrix is F16 and tile-aligned
 * (default 16; 0 = always replay the per-token decode step like the reference's forward loop) */
void nt_model_set_prefill_min_tokens(nt_model_t m, int n);
/* algorithmic bytes read per decoded token by this rank at context length ctx */
unsigned long long nt_model_bytes_per_token(nt_model_t m, int ctx);

/* Samples the next token from the last logits ON THE GPU: repeat penalty over recent_window (host ids, oldest first; applied in
 * place to the device logits), temperature, top-k, top-p and the inverse-CDF draw with the uniform variate r — the steps of the
 * reference's Sampler::apply_repeat_penalty + Sampler::sample (src/inference/sampler.cpp:30-117) in the same order.  Returns the
 * token id, or -1 when the settings are not covered (temperature <= 0, top_k <= 0, top_k > 1024, top_k >= vocab): sample on the
 * host then (nt_sample_token).  nt_sampler_uniform(seed, i) is the (i+1)-th variate of the reference's mt19937 stream. */
int   nt_model_sample(nt_model_t m, float temperature, int top_k, float top_p, float repeat_penalty, const int* recent_window,
                      int n_window, float r);
float nt_sampler_uniform(uint64_t seed, int n_draws_before);

/* Opt-in: run each decoded token as ONE persistent kernel (all layers; grid barriers and, under tensor parallelism,
 * NVLink peer-memory exchanges inside the kernel) instead of the CUDA graph of fused launches.  Same maths, same KV
 * cache.  Models whose shapes/dtypes the kernel does not cover keep the graph path (a note goes to stderr).
 * Also enabled by the environment variable NT_B200_MEGAKERNEL=1.  Call before the first forward. */
void nt_model_use_megakernel(nt_model_t m, int on);
/* wall-clock seconds nt_model_load_gguf took for this model (header parse + shard plan + pipelined pread -> pinned staging ->
 * cudaMemcpyAsync upload + buffer allocation); 0 for models built from device tensors.  Reference behaviour it replaces: one
 * synchronous cudaMemcpy per tensor from the pageable mmap (src/model/transformer.cpp:286-328, src/core/tensor.cpp:224-225). */
double nt_model_load_seconds(nt_model_t m);
/* how a tensor-parallel model sums the o-projection / down-projection partials: 0 not tensor parallel (or no step run yet),
 * 1 ncclAllReduce, 2 NVLink peer-memory exchange fused into the GEMV epilogue (default; NT_B200_TP_NCCL=1 selects 1) */
int  nt_model_tp_exchange(nt_model_t m);
int  nt_model_megakernel_active(nt_model_t m);      /* 1 once the persistent kernel has been built and is in use */
/* phase kinds of the persistent kernel's per-token program (0 norm+quantise, 1 quantise, 2 GEMV, 3 attention,
 * 4 combine); returns the number of phases (may exceed cap), 0 when the kernel is not active */
int  nt_model_megakernel_plan(nt_model_t m, int* kinds, int cap);
/* tuning aid: record the SM clock at every phase boundary of the persistent kernel on its first 4 CTAs; _read returns the number of
 * values (per CTA: [phases][start, work done, barrier passed] ticks, then clock start/end and %globaltimer start/end for the
 * tick -> ns conversion) of the most recent launch and copies up to cap of them */
void nt_model_megakernel_trace(nt_model_t m, int on);
long long nt_model_megakernel_trace_read(nt_model_t m, unsigned long long* out_host, size_t cap);
/* debug: copies one of the persistent kernel's working vectors ("hid0", "hid1", "q", "attn", "act", "slots") to the host;
 * returns its length in floats, -1 when unknown / not active */
long long nt_model_debug_read(nt_model_t m, const char* name, float* out_host, size_t cap);

/* Host-only self-test of the persistent kernel's plan builder: builds the per-token program for one tensor-parallel rank
 * of a model of shape *cfg (layer_dtypes: n_layers x 7 nt::DType values in the order q,k,v,o,gate,up,down) on a grid of
 * `grid` CTAs (fuse: producer-side fusion bits, 1 = activation quantiser, 2 = split combine, 4 = residual add + next norm, single rank only) and replays the schedule of every GEMV phase on the CPU.  info (8 ints, optional): phases, body phases, GEMV
 * phases, min warps, min ring stages, exchanges, attention splits, keys per split.  Returns 0 ok, 1 shape not covered,
 * 2 inconsistent plan (a bug), -1 bad arguments; msg receives the reason. */
int  nt_mega_plan_selftest(const nt_model_config* cfg, int tp_rank, int tp_size, const int* layer_dtypes, int head_dtype,
                           int grid, int split_fixed, int fuse, int* info, char* msg, size_t cap);

/* ---- host-only helpers (no GPU needed): GGUF header, tokenizer and sampler of the CLI path ---- */
/* JSON description of a GGUF file as the engine parses it (config, vocab size, tensor table). Returns the
 * number of bytes written (excluding NUL) or -1. */
int  nt_gguf_describe(const char* path, char* out, size_t cap);
int  nt_tokenize(const char* gguf_path, const char* text, int add_bos, int* ids, int cap);      /* returns count or -1 */
int  nt_detokenize(const char* gguf_path, const int* ids, int n, char* out, size_t cap);       /* returns bytes or -1 */
/* One draw of the reference-compatible sampler (repeat penalty over `recent`, temperature, top-k, top-p, mt19937(seed)). */
int  nt_sample_token(const float* logits, int n, float temperature, int top_k, float top_p, float repeat_penalty,
                     int repeat_window, const int* recent, int n_recent, uint64_t seed);

/* tensor-parallel communicator */
int  nt_tp_unique_id(void* out128);                 /* rank 0; returns 0 on success */
int  nt_tp_init(nt_model_t m, const void* id128, int rank, int size);

#ifdef __cplusplus
}
#endif
#endif
