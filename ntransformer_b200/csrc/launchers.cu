// launchers.cu — the drop-in boundary.
//   (1) nt::cuda::launch_*  : the reference's C++ launcher names and signatures (src/cuda/kernels.h:10-74),
//       exported with the same Itanium mangling so reference model/*.cpp and tests/test_gemm.cpp link unchanged;
//   (2) extern "C" nt_b200_* : the same entry points as a true C-ABI (include/nt_b200.h);
//   (3) extern "C" nt_cuda_* : the reference's memory helpers (src/core/device.h:79-88, device.cu:152-198).
// Launchers are fire-and-forget like the reference's (no error returned, unsupported dtype -> stderr + no-op).
#include "kernels_internal.h"
#include "../../include/nt_b200.h"

namespace nt { namespace b200 {

static void gemv_dispatch(float* y, const void* W, const float* x, int out, int in, DType dt, GemvEpilogue ep, cudaStream_t s) {
    if (out <= 0 || in <= 0) return;
    GemvMat m;
    m.W = W; m.y = y; m.out = out; m.dtype = dt; m.row_pitch = 0;
    if ((dt == DType::Q4_K_M || dt == DType::Q5_K || dt == DType::Q6_K || dt == DType::Q8_0) && gemv_kq_supported(&m, 1, in)) {
        GemvInput gi;                 // F32 activations are quantised inside the kernel prologue: one launch, no state
        gi.x = x;
        gemv_kq(&m, 1, in, gi, ep, s);
    } else {
        gemv_generic(y, W, x, out, in, dt, 0, ep, s);
    }
}

}}  // namespace nt::b200

// ------------------------------------------------------------------------------------------------
// (1) reference C++ launcher names
// ------------------------------------------------------------------------------------------------
namespace nt { namespace cuda {
using namespace nt::b200;
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

void launch_rmsnorm(float* output, const float* input, const float* weight, int batch_size, int hidden_size, float eps, void* stream) {
    rmsnorm(output, input, weight, batch_size, hidden_size, eps, S(stream));
}
void launch_rmsnorm_f16(void* output, const float* input, const float* weight, int batch_size, int hidden_size, float eps, void* stream) {
    rmsnorm_f16(output, input, weight, batch_size, hidden_size, eps, S(stream));
}
void launch_rope(float* q, float* k, const int* positions, int batch_size, int seq_len, int n_heads, int n_kv_heads,
                 int head_dim, float theta_base, float freq_scale, bool interleaved, void* stream) {
    (void)batch_size;   // unused by the reference too (rotary.cu:116)
    rope(q, k, positions, seq_len, n_heads, n_kv_heads, head_dim, theta_base, freq_scale, interleaved, S(stream));
}
void launch_softmax(float* output, const float* input, int rows, int cols, void* stream) { softmax(output, input, rows, cols, S(stream)); }
void launch_masked_softmax(float* output, const float* input, const bool* mask, int rows, int cols, void* stream) {
    masked_softmax(output, input, mask, rows, cols, S(stream));
}
void launch_gemv(float* y, const void* W, const float* x, int out_features, int in_features, DType weight_dtype, void* stream) {
    gemv_dispatch(y, W, x, out_features, in_features, weight_dtype, GEMV_STORE, S(stream));
}
void launch_gemv_add(float* y, const void* W, const float* x, int out_features, int in_features, DType weight_dtype, void* stream) {
    if (weight_dtype != DType::F16) {      // gemm.cu:866-869
        fprintf(stderr, "launch_gemv_add: only F16 supported (got %s)\n", dtype_name(weight_dtype));
        return;
    }
    gemv_generic(y, W, x, out_features, in_features, weight_dtype, 0, GEMV_ADD, S(stream));
}
void launch_gemm_f32(float* C, const float* A, const float* B, int M, int N, int K, void* stream) { gemm_f32(C, A, B, M, N, K, S(stream)); }
void launch_silu_mul(float* output, const float* gate, const float* up, int size, void* stream) { silu_mul(output, gate, up, size, S(stream)); }
void launch_add_bias(float* y, const float* bias, int size, void* stream) { add_bias(y, bias, size, S(stream)); }
void launch_attention_decode(float* output, const float* q, const void* k_cache, const void* v_cache, int seq_len, int n_heads,
                             int n_kv_heads, int head_dim, int max_seq, float scale, void* stream) {
    attention_decode(output, q, k_cache, v_cache, seq_len, n_heads, n_kv_heads, head_dim, max_seq, scale, S(stream));
}
void launch_attention_prefill(float* output, const float* Q, const void* k_cache, const void* v_cache, int seq_len, int start_pos,
                              int n_heads, int n_kv_heads, int head_dim, int max_seq, float scale, void* stream) {
    attention_prefill(output, Q, k_cache, v_cache, seq_len, start_pos, n_heads, n_kv_heads, head_dim, max_seq, scale, S(stream));
}
void launch_copy_to_kv_cache(void* k_cache, void* v_cache, const float* k, const float* v, int seq_len, int n_kv_heads,
                             int head_dim, int start_pos, int max_seq, void* stream) {
    copy_to_kv_cache(k_cache, v_cache, k, v, seq_len, n_kv_heads, head_dim, start_pos, max_seq, S(stream));
}
void launch_add(float* out, const float* a, const float* b, int size, void* stream) { add(out, a, b, size, S(stream)); }
void launch_add_inplace(float* a, const float* b, int size, void* stream) { add_inplace(a, b, size, S(stream)); }
void launch_copy(float* dst, const float* src, int size, void* stream) { copy(dst, src, size, S(stream)); }
void launch_cosine_similarity(float* result, const float* a, const float* b, int size, void* stream) {
    cosine_similarity(result, a, b, size, S(stream));
}
}}  // namespace nt::cuda

// ------------------------------------------------------------------------------------------------
// (2) C-ABI
// ------------------------------------------------------------------------------------------------
using nt::DType;
namespace C = nt::cuda;
extern "C" {
void nt_b200_rmsnorm(float* o, const float* i, const float* w, int b, int h, float eps, void* s) { C::launch_rmsnorm(o, i, w, b, h, eps, s); }
void nt_b200_rmsnorm_f16(void* o, const float* i, const float* w, int b, int h, float eps, void* s) { C::launch_rmsnorm_f16(o, i, w, b, h, eps, s); }
void nt_b200_rope(float* q, float* k, const int* pos, int b, int sl, int nh, int nkv, int hd, float th, float fs, int il, void* s) {
    C::launch_rope(q, k, pos, b, sl, nh, nkv, hd, th, fs, il != 0, s);
}
void nt_b200_softmax(float* o, const float* i, int r, int c, void* s) { C::launch_softmax(o, i, r, c, s); }
void nt_b200_masked_softmax(float* o, const float* i, const unsigned char* m, int r, int c, void* s) {
    C::launch_masked_softmax(o, i, reinterpret_cast<const bool*>(m), r, c, s);
}
void nt_b200_gemv(float* y, const void* W, const float* x, int out, int in, int dt, void* s) { C::launch_gemv(y, W, x, out, in, (DType)dt, s); }
void nt_b200_gemv_add(float* y, const void* W, const float* x, int out, int in, int dt, void* s) { C::launch_gemv_add(y, W, x, out, in, (DType)dt, s); }
void nt_b200_gemm_f32(float* Cm, const float* A, const float* B, int M, int N, int K, void* s) { C::launch_gemm_f32(Cm, A, B, M, N, K, s); }
void nt_b200_silu_mul(float* o, const float* g, const float* u, int n, void* s) { C::launch_silu_mul(o, g, u, n, s); }
void nt_b200_add_bias(float* y, const float* b, int n, void* s) { C::launch_add_bias(y, b, n, s); }
void nt_b200_attention_decode(float* o, const float* q, const void* kc, const void* vc, int sl, int nh, int nkv, int hd, int ms, float sc, void* s) {
    C::launch_attention_decode(o, q, kc, vc, sl, nh, nkv, hd, ms, sc, s);
}
void nt_b200_attention_prefill(float* o, const float* Q, const void* kc, const void* vc, int sl, int sp, int nh, int nkv, int hd, int ms, float sc, void* s) {
    C::launch_attention_prefill(o, Q, kc, vc, sl, sp, nh, nkv, hd, ms, sc, s);
}
void nt_b200_copy_to_kv_cache(void* kc, void* vc, const float* k, const float* v, int sl, int nkv, int hd, int sp, int ms, void* s) {
    C::launch_copy_to_kv_cache(kc, vc, k, v, sl, nkv, hd, sp, ms, s);
}
void nt_b200_add(float* o, const float* a, const float* b, int n, void* s) { C::launch_add(o, a, b, n, s); }
void nt_b200_add_inplace(float* a, const float* b, int n, void* s) { C::launch_add_inplace(a, b, n, s); }
void nt_b200_copy(float* d, const float* sr, int n, void* s) { C::launch_copy(d, sr, n, s); }
void nt_b200_cosine_similarity(float* r, const float* a, const float* b, int n, void* s) { C::launch_cosine_similarity(r, a, b, n, s); }

size_t nt_b200_xq_bytes(int K) { return nt::b200::xq_bytes(K); }
void nt_b200_quantize_x(const float* x, void* xq, int K, void* s) { nt::b200::quantize_x(x, xq, K, static_cast<cudaStream_t>(s)); }
int nt_b200_gemv_fused(int n_mat, float* const* y, const void* const* W, const int* out_features, const int* dtypes,
                       int in_features, const void* xq, int epilogue, void* stream) {
    if (n_mat < 1 || n_mat > 3) return -1;
    nt::b200::GemvMat m[3];
    for (int i = 0; i < n_mat; i++) { m[i].W = W[i]; m[i].y = y[i]; m[i].out = out_features[i]; m[i].dtype = (DType)dtypes[i]; m[i].row_pitch = 0; }
    if (!nt::b200::gemv_kq_supported(m, n_mat, in_features)) return -2;
    if (epilogue == 2 && (n_mat != 2 || out_features[0] != out_features[1])) return -3;
    nt::b200::gemv_kq(m, n_mat, in_features, xq, (nt::b200::GemvEpilogue)epilogue, static_cast<cudaStream_t>(stream));
    return 0;
}
int nt_b200_gemv_fused_f32(int n_mat, float* const* y, const void* const* W, const int* out_features, const int* dtypes, int in_features,
                           const float* x, const float* norm_w, float eps, int epilogue, void* stream) {
    if (n_mat < 1 || n_mat > 3 || !x) return -1;
    nt::b200::GemvMat m[3];
    for (int i = 0; i < n_mat; i++) { m[i].W = W[i]; m[i].y = y[i]; m[i].out = out_features[i]; m[i].dtype = (DType)dtypes[i]; m[i].row_pitch = 0; }
    if (!nt::b200::gemv_kq_supported(m, n_mat, in_features)) return -2;
    if (epilogue == 2 && (n_mat != 2 || out_features[0] != out_features[1])) return -3;
    nt::b200::GemvInput in; in.x = x; in.norm_w = norm_w; in.eps = eps;
    nt::b200::gemv_kq(m, n_mat, in_features, in, (nt::b200::GemvEpilogue)epilogue, static_cast<cudaStream_t>(stream));
    return 0;
}
size_t nt_b200_attention_decode_scratch_floats(int ms, int nh, int nkv, int hd) { return nt::b200::attention_decode_dyn_scratch_floats(ms, nh, nkv, hd); }
int nt_b200_attention_decode_tickets(int nh, int nkv) { return nt::b200::attention_decode_fused_tickets(nh, nkv); }
void nt_b200_attention_decode_fused(float* o, const float* q, const float* k, const float* v, void* kc, void* vc, const int* pos_dev, int ms,
                                    int nh, int nkv, int hd, float theta, float fs, float sc, float* scratch, unsigned* tickets, void* xq_out,
                                    void* s) {
    nt::b200::attention_decode_fused(o, q, k, v, kc, vc, pos_dev, ms, nh, nkv, hd, theta, fs, sc, scratch, tickets, xq_out,
                                     static_cast<cudaStream_t>(s));
}
void nt_b200_embed_rows(float* out, const void* table, int dt, const int* tokens_dev, int n, int hidden, void* s) {
    nt::b200::embed_rows(out, table, (DType)dt, tokens_dev, n, hidden, static_cast<cudaStream_t>(s));
}
size_t nt_b200_gemm_f16_tc_workspace_bytes(int M, int K) { return nt::b200::gemm_f16_tc_workspace_bytes(M, K); }
int nt_b200_gemm_f16_tc(float* Cm, const float* A, const void* W, int M, int N, int K, void* ws, void* s) {
    return nt::b200::gemm_f16_tc(Cm, A, W, M, N, K, ws, static_cast<cudaStream_t>(s)) ? 0 : -1;
}
void nt_b200_split_activations(void* ws, const float* A, int M, int K, void* s) {
    nt::b200::split_activations(ws, A, M, K, static_cast<cudaStream_t>(s));
}
int nt_b200_gemm_f16_tc_ws(float* Cm, const void* ws, const void* W, int M, int N, int K, int add, void* s) {
    return nt::b200::gemm_f16_tc_ws(Cm, ws, W, M, N, K, add != 0, static_cast<cudaStream_t>(s)) ? 0 : -1;
}
int nt_b200_gemm_f16_tc_swiglu_ws(void* wo, const void* wi, const void* Wg, const void* Wu, int M, int N, int K, void* s) {
    return nt::b200::gemm_f16_tc_swiglu_ws(wo, wi, Wg, Wu, M, N, K, static_cast<cudaStream_t>(s)) ? 0 : -1;
}
void nt_b200_rmsnorm_split(void* ws, const float* x, const float* w, int rows, int hidden, float eps, void* s) {
    nt::b200::rmsnorm_split(ws, x, w, rows, hidden, eps, static_cast<cudaStream_t>(s));
}
void nt_b200_dequant_split(void* hi, void* lo, const void* W, int dt, size_t pitch, int rows, int cols, void* s) {
    nt::b200::dequant_split(hi, lo, W, (DType)dt, pitch, rows, cols, static_cast<cudaStream_t>(s));
}
unsigned long long nt_b200_launch_count(void) { return nt::b200::launch_count(); }
int nt_b200_stream_sync(void* s) { return (int)cudaStreamSynchronize(static_cast<cudaStream_t>(s)); }
const char* nt_b200_version(void) { return "ntransformer_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// (3) memory helpers — abort on failure like the reference's NT_CUDA_CHECK (device.cu:152-198)
// ------------------------------------------------------------------------------------------------
void* nt_cuda_malloc(size_t size) { void* p = nullptr; NT_CUDA_CHECK(cudaMalloc(&p, size)); return p; }
void nt_cuda_free(void* ptr) { if (ptr) NT_CUDA_CHECK(cudaFree(ptr)); }
void nt_cuda_memcpy_h2d(void* dst, const void* src, size_t size) { NT_CUDA_CHECK(cudaMemcpy(dst, src, size, cudaMemcpyHostToDevice)); }
void nt_cuda_memcpy_d2h(void* dst, const void* src, size_t size) { NT_CUDA_CHECK(cudaMemcpy(dst, src, size, cudaMemcpyDeviceToHost)); }
void nt_cuda_memcpy_d2d(void* dst, const void* src, size_t size) { NT_CUDA_CHECK(cudaMemcpy(dst, src, size, cudaMemcpyDeviceToDevice)); }
void nt_cuda_memset(void* ptr, int value, size_t size) { NT_CUDA_CHECK(cudaMemset(ptr, value, size)); }
void* nt_cuda_malloc_host(size_t size) { void* p = nullptr; NT_CUDA_CHECK(cudaMallocHost(&p, size)); return p; }
void nt_cuda_free_host(void* ptr) { if (ptr) NT_CUDA_CHECK(cudaFreeHost(ptr)); }
}  // extern "C"
