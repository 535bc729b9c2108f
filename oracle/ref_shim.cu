// ref_shim.cu — OUR thin C wrappers around the UNMODIFIED reference (compiled from /root/reference where
// it lies, see oracle/Makefile `ref`).  Test infrastructure: gives tests and bench.py --impl reference
// a C-ABI onto the reference's own CUDA kernels (nt::cuda::launch_*, src/cuda/kernels.h:10-74) and its
// own resident forward (nt::Transformer::forward, src/model/transformer.cpp:604-669) on the same box.
// Linked with -Bsymbolic so the reference's symbols inside this .so never bind to libnt_b200.so's.
#include "src/cuda/kernels.h"
#include "src/core/device.h"
#include "src/model/transformer.h"
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

}  // extern "C"
