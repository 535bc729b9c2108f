/*
 * nt_b200.h — C-ABI of the B200-native resident-decode kernels (libnt_b200.so).
 *
 * Drop-in boundary: every entry point below replaces one launcher of the reference's
 * kernel interface  /root/reference/src/cuda/kernels.h:10-74  (C++ free functions in
 * namespace nt::cuda) or one of its extern "C" memory helpers
 * /root/reference/src/core/device.h:79-88.  Same argument order, meaning, units and
 * (absence of) error reporting as the reference; `dtype` carries the numeric value of
 * nt::DType (reference src/core/types.h:24-35: F32 0, F16 1, Q8_0 2, Q4_0 3, Q4_K_M 4,
 * Q6_K 5, Q5_K 6); `stream` is a cudaStream_t passed as void*.  All pointers are DEVICE
 * pointers owned by the caller; launches are asynchronous on `stream`.  The library also
 * exports the original mangled names (nt::cuda::launch_*) so the unmodified reference host
 * code links against it — see INTEGRATION.md.
 *
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef NT_B200_H
#define NT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- kernels.h:14-19  launch_rmsnorm / launch_rmsnorm_f16 ---- */
void nt_b200_rmsnorm(float* output, const float* input, const float* weight,
                     int batch_size, int hidden_size, float eps, void* stream);
void nt_b200_rmsnorm_f16(void* output, const float* input, const float* weight,
                         int batch_size, int hidden_size, float eps, void* stream);
/* ---- kernels.h:22-25  launch_rope ---- */
void nt_b200_rope(float* q, float* k, const int* positions, int batch_size, int seq_len,
                  int n_heads, int n_kv_heads, int head_dim, float theta_base,
                  float freq_scale, int interleaved, void* stream);
/* ---- kernels.h:28-31  launch_softmax / launch_masked_softmax ---- */
void nt_b200_softmax(float* output, const float* input, int rows, int cols, void* stream);
void nt_b200_masked_softmax(float* output, const float* input, const unsigned char* mask,
                            int rows, int cols, void* stream);
/* ---- kernels.h:34-36  launch_gemv : y[out] = W[out,in] . x, W in GGUF block layout ---- */
void nt_b200_gemv(float* y, const void* W, const float* x, int out_features,
                  int in_features, int weight_dtype, void* stream);
/* ---- kernels.h:39-41  launch_gemv_add : y += W . x (F16 only, like the reference) ---- */
void nt_b200_gemv_add(float* y, const void* W, const float* x, int out_features,
                      int in_features, int weight_dtype, void* stream);
/* ---- kernels.h:42-44  launch_gemm_f32 : C[M,N] = A[M,K] . B[N,K]^T ---- */
void nt_b200_gemm_f32(float* C, const float* A, const float* B, int M, int N, int K, void* stream);
/* ---- kernels.h:45-48  launch_silu_mul / launch_add_bias ---- */
void nt_b200_silu_mul(float* output, const float* gate, const float* up, int size, void* stream);
void nt_b200_add_bias(float* y, const float* bias, int size, void* stream);
/* ---- kernels.h:51-63  attention over the F16 KV cache [max_seq, n_kv_heads, head_dim] ---- */
void nt_b200_attention_decode(float* output, const float* q, const void* k_cache,
                              const void* v_cache, int seq_len, int n_heads, int n_kv_heads,
                              int head_dim, int max_seq, float scale, void* stream);
void nt_b200_attention_prefill(float* output, const float* Q, const void* k_cache,
                               const void* v_cache, int seq_len, int start_pos, int n_heads,
                               int n_kv_heads, int head_dim, int max_seq, float scale, void* stream);
void nt_b200_copy_to_kv_cache(void* k_cache, void* v_cache, const float* k, const float* v,
                              int seq_len, int n_kv_heads, int head_dim, int start_pos,
                              int max_seq, void* stream);
/* ---- kernels.h:66-71  element-wise ops and cosine similarity ---- */
void nt_b200_add(float* out, const float* a, const float* b, int size, void* stream);
void nt_b200_add_inplace(float* a, const float* b, int size, void* stream);
void nt_b200_copy(float* dst, const float* src, int size, void* stream);
void nt_b200_cosine_similarity(float* result, const float* a, const float* b, int size, void* stream);

/* ---- core/device.h:79-88  memory helpers (same names as the reference exports) ---- */
void* nt_cuda_malloc(size_t size);
void  nt_cuda_free(void* ptr);
void  nt_cuda_memcpy_h2d(void* dst, const void* src, size_t size);
void  nt_cuda_memcpy_d2h(void* dst, const void* src, size_t size);
void  nt_cuda_memcpy_d2d(void* dst, const void* src, size_t size);
void  nt_cuda_memset(void* ptr, int value, size_t size);
void* nt_cuda_malloc_host(size_t size);
void  nt_cuda_free_host(void* ptr);

/* ---- additions (no reference counterpart): fused-path building blocks and introspection ---- */
/* bytes of the block-scaled int8x3 activation buffer for K elements */
size_t nt_b200_xq_bytes(int K);
void   nt_b200_quantize_x(const float* x, void* xq, int K, void* stream);
/* fused K-quant GEMV over n_mat (<=3) matrices sharing one quantised activation vector.
 * epilogue: 0 store, 1 y += W.x, 2 y0 = silu(W0.x) * (W1.x) */
int    nt_b200_gemv_fused(int n_mat, float* const* y, const void* const* W, const int* out_features,
                          const int* dtypes, int in_features, const void* xq, int epilogue, void* stream);
/* nt_b200_gemv_fused fed with the F32 activation vector itself: the kernel quantises x in its prologue (one pass, staged over
 * the part of its TMA ring that is primed afterwards), so no separate quantise launch is needed.  norm_w (may be NULL): the
 * GEMV of RMSNorm(x) * norm_w (src/cuda/rmsnorm.cu:17-70 followed by launch_gemv): the prologue quantises x * norm_w, sums x^2
 * in the same pass and scales the results by rsqrt(mean(x^2) + eps) — equal to normalise-then-multiply up to round-off.
 * Replaces the launch pairs launch_rmsnorm + launch_gemv of src/model/attention.cpp:144-162 and src/model/ffn.cpp:96-133.
 * Returns 0, or < 0 when rejected (nothing launched). */
int    nt_b200_gemv_fused_f32(int n_mat, float* const* y, const void* const* W, const int* out_features,
                              const int* dtypes, int in_features, const float* x, const float* norm_w, float eps,
                              int epilogue, void* stream);
/* The decode step's attention sub-block in ONE launch: launch_rope (src/cuda/rotary.cu:113-140; q and k are read, not
 * modified) + launch_copy_to_kv_cache at row *pos_dev (src/cuda/attention.cu:405-425) + launch_attention_decode over
 * *pos_dev + 1 keys (src/cuda/attention.cu:348-375), context split over CTAs, merged by the last CTA of each head group.
 * scratch: nt_b200_attention_decode_scratch_floats() floats; tickets: nt_b200_attention_decode_tickets() zeroed words (left
 * zeroed); xq_out (may be NULL): the output also in xq form.  head_dim 64 / 128 / 256. */
size_t nt_b200_attention_decode_scratch_floats(int max_seq, int n_heads, int n_kv_heads, int head_dim);
int    nt_b200_attention_decode_tickets(int n_heads, int n_kv_heads);
void   nt_b200_attention_decode_fused(float* out, const float* q, const float* k, const float* v, void* k_cache,
                                      void* v_cache, const int* pos_dev, int max_seq, int n_heads, int n_kv_heads,
                                      int head_dim, float theta_base, float freq_scale, float scale, float* scratch,
                                      unsigned* tickets, void* xq_out, void* stream);
void   nt_b200_embed_rows(float* out, const void* table, int dtype, const int* tokens_dev,
                          int n_tokens, int hidden, void* stream);
/* Prefill GEMM on the tcgen05 tensor cores: C[M,N] (F32, row-major) = A[M,K] (F32) . W[N,K]^T (F16 weights, GGUF F16
 * row-major).  The reference has no live counterpart (its forward loops launch_gemv per token: src/model/attention.cpp:144-162;
 * launch_gemm_f32, src/cuda/kernels.h:72-74, is unused).  N % 128 == 0 and K % 64 == 0; workspace holds the F16 hi/lo split
 * of A.  Returns 0, or -1 for unsupported shapes (nothing launched). */
size_t nt_b200_gemm_f16_tc_workspace_bytes(int M, int K);
int    nt_b200_gemm_f16_tc(float* C, const float* A, const void* W_f16, int M, int N, int K,
                           void* workspace, void* stream);
/* the two halves of nt_b200_gemm_f16_tc, for callers that reuse one split of A across several weight matrices
 * (q/k/v, gate/up); add != 0 accumulates into C (residual epilogue) */
void   nt_b200_split_activations(void* workspace, const float* A, int M, int K, void* stream);
int    nt_b200_gemm_f16_tc_ws(float* C, const void* workspace, const void* W_f16, int M, int N, int K,
                              int add, void* stream);
/* FFN front half for prefill: workspace_out = split(silu(A.Wgate^T) * (A.Wup^T)) with A pre-split in workspace_in
 * (reference: ffn.cpp:96-133, three launches per token); N = rows of Wgate/Wup (% 128), feeds nt_b200_gemm_f16_tc_ws */
int    nt_b200_gemm_f16_tc_swiglu_ws(void* workspace_out, const void* workspace_in, const void* Wgate_f16,
                                     const void* Wup_f16, int M, int N, int K, void* stream);
/* RMSNorm (launch_rmsnorm math) written straight into the split workspace */
void   nt_b200_rmsnorm_split(void* workspace, const float* x, const float* w, int rows, int hidden, float eps,
                             void* stream);
/* Any GGUF weight matrix (dtype = numeric nt::DType, row_pitch 0 = dense rows) -> dense F16 pair with W = w_hi + w_lo, the
 * operands of two nt_b200_gemm_f16_tc_ws launches (second with add = 1): tensor-core prefill for quantised models.
 * Reference dequant formulas: src/cuda/gemm.cu:32-470, src/model/transformer.cpp:394-599. */
void   nt_b200_dequant_split(void* w_hi, void* w_lo, const void* W, int dtype, size_t row_pitch, int rows, int cols,
                             void* stream);
unsigned long long nt_b200_launch_count(void);   /* kernels launched by this library so far */
int    nt_b200_stream_sync(void* stream);        /* cudaStreamSynchronize; returns cudaError_t */
const char* nt_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NT_B200_H */
