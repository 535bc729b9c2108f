// ref_shim.cu — OUR thin C wrappers around the UNMODIFIED reference (compiled from /root/reference where
// it lies, see oracle/Makefile `ref`).  Test infrastructure: gives tests and bench.py --impl reference
// a C-ABI onto the reference's own CUDA kernels (nt::cuda::launch_*, src/cuda/kernels.h:10-74) and its
// own resident forward (nt::Transformer::forward, src/model/transformer.cpp:604-669) on the same box.
// Linked with -Bsymbolic so the reference's symbols inside this .so never bind to libnt_b200.so's.
#include "src/cuda/kernels.h"
#include "src/core/device.h"
#include "src/model/transformer.h"
#include "src/model/loader.h"
#include "src/inference/tokenizer.h"
#include "src/inference/sampler.h"
#include <cstring>
#include <cuda_runtime.h>
#include <chrono>
#include <memory>
#include <string>
#include <vector>

using nt::DType;
namespace rc = nt::cuda;

extern "C" {

void ref_gemv(float* y, const void* W, const float* x, int out, int in, int dt, void* s) { rc::launch_gemv(y, W, x, out, in, (DType)dt, s); }
void ref_rmsnorm(float* o, const float* i, const float* w, int b, int h, float eps, void* s) { rc::launch_rmsnorm(o, i, w, b, h, eps, s); }
void ref_rope(float* q, float* k, const int* pos, int b, int sl, int nh, int nkv, int hd, float th, float fs, int il, void* s) {
    rc::launch_rope(q, k, pos, b, sl, nh, nkv, hd, th, fs, il != 0, s);
}
void ref_silu_mul(float* o, const float* g, const float* u, int n, void* s) { rc::launch_silu_mul(o, g, u, n, s); }
void ref_add_inplace(float* a, const float* b, int n, void* s) { rc::launch_add_inplace(a, b, n, s); }
void ref_attention_decode(float* o, const float* q, const void* kc, const void* vc, int sl, int nh, int nkv, int hd, int ms, float sc, void* s) {
    rc::launch_attention_decode(o, q, kc, vc, sl, nh, nkv, hd, ms, sc, s);
}
void ref_attention_prefill(float* o, const float* Q, const void* kc, const void* vc, int sl, int sp, int nh, int nkv, int hd, int ms, float sc, void* s) {
    rc::launch_attention_prefill(o, Q, kc, vc, sl, sp, nh, nkv, hd, ms, sc, s);
}
void ref_copy_to_kv_cache(void* kc, void* vc, const float* k, const float* v, int sl, int nkv, int hd, int sp, int ms, void* s) {
    rc::launch_copy_to_kv_cache(kc, vc, k, v, sl, nkv, hd, sp, ms, s);
}

// ---- whole-model probe: token ids in, logits out, through the reference's own Transformer ----
struct RefModel { nt::Transformer tf; };

void* ref_model_load(const char* gguf_path, int max_ctx) {
    auto* m = new RefModel();
    if (!m->tf.load(gguf_path, max_ctx, /*streaming=*/false)) { delete m; return nullptr; }
    return m;
}
void ref_model_free(void* h) { delete static_cast<RefModel*>(h); }
int ref_model_vocab(void* h) { return static_cast<RefModel*>(h)->tf.config().vocab_size; }
// Runs forward(tokens, n, start_pos); copies vocab logits to host; returns device+sync time in ms.
float ref_model_forward(void* h, const int* tokens, int n, int start_pos, float* logits_host) {
    auto* m = static_cast<RefModel*>(h);
    auto t0 = std::chrono::steady_clock::now();
    float* dl = m->tf.forward(tokens, n, start_pos);   // synchronises STREAM_COMPUTE itself (transformer.cpp:667)
    auto t1 = std::chrono::steady_clock::now();
    if (logits_host) cudaMemcpy(logits_host, dl, sizeof(float) * (size_t)m->tf.config().vocab_size, cudaMemcpyDeviceToHost);
    return std::chrono::duration<float, std::milli>(t1 - t0).count();
}

// ---- host-only parts of the path (no GPU needed): the reference's GGUF parser, tokenizer and sampler ----
// Same signatures as nt_gguf_describe-lite / nt_tokenize / nt_detokenize / nt_sample_token of include/nt_b200_engine.h,
// so tests/test_ref_host.py can put the two implementations side by side on the same file and inputs.

// out[0..11] = vocab, hidden, inter, layers, heads, kv heads, head_dim, max_seq, bos, eos, n_vocab_tokens, data_offset;
// fout[0..1] = norm_eps, rope_theta.  Returns the tensor count or -1.
int ref_gguf_config(const char* path, long long* out, float* fout) {
    nt::GGUFLoader L;
    if (!L.load(path)) return -1;
    const nt::ModelConfig& c = L.config();
    out[0] = c.vocab_size; out[1] = c.hidden_size; out[2] = c.intermediate_size; out[3] = c.n_layers; out[4] = c.n_heads;
    out[5] = c.n_kv_heads; out[6] = c.head_dim; out[7] = c.max_seq_len; out[8] = c.bos_token_id; out[9] = c.eos_token_id;
    out[10] = (long long)L.vocab().tokens.size(); out[11] = (long long)L.data_offset();
    fout[0] = c.norm_eps; fout[1] = c.rope_theta;
    return (int)L.tensor_names().size();
}
// info[0..2] = dtype (nt::DType of the reference), offset inside the data section, nbytes; returns 0 or -1.
int ref_gguf_tensor(const char* path, const char* name, long long* info) {
    nt::GGUFLoader L;
    if (!L.load(path)) return -1;
    const nt::GGUFTensorInfo* t = L.tensor_info(name);
    if (!t) return -1;
    info[0] = (long long)nt::ggml_to_dtype(t->ggml_type); info[1] = (long long)t->offset; info[2] = (long long)t->nbytes;
    return 0;
}
int ref_tokenize(const char* gguf_path, const char* text, int add_bos, int* ids, int cap) {
    nt::GGUFLoader L;
    if (!L.load(gguf_path)) return -1;
    nt::Tokenizer t;
    t.init(L.vocab(), L.config().bos_token_id, L.config().eos_token_id);
    std::vector<int> v = t.encode(text, add_bos != 0);
    for (int i = 0; i < (int)v.size() && i < cap; i++) ids[i] = v[(size_t)i];
    return (int)v.size();
}
int ref_detokenize(const char* gguf_path, const int* ids, int n, char* out, size_t cap) {
    nt::GGUFLoader L;
    if (!L.load(gguf_path)) return -1;
    nt::Tokenizer t;
    t.init(L.vocab(), L.config().bos_token_id, L.config().eos_token_id);
    std::string s = t.decode(std::vector<int>(ids, ids + n));
    if (s.size() + 1 > cap) return -1;
    memcpy(out, s.data(), s.size());
    out[s.size()] = 0;
    return (int)s.size();
}
int ref_sample_token(const float* logits, int n, float temperature, int top_k, float top_p, float repeat_penalty,
                     int repeat_window, const int* recent, int n_recent, unsigned long long seed) {
    nt::SamplerConfig sc;
    sc.temperature = temperature; sc.top_k = top_k; sc.top_p = top_p; sc.repeat_penalty = repeat_penalty;
    sc.repeat_window = repeat_window; sc.seed = seed;
    nt::Sampler s;
    s.init(sc);
    std::vector<float> l(logits, logits + n);
    s.apply_repeat_penalty(l.data(), n, std::vector<int>(recent, recent + (recent ? n_recent : 0)));
    return s.sample(l.data(), n);
}

}  // extern "C"
