/*
 * nt_oracle.h — CPU restatement of the reference's resident decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ntransformer_b200/ may include,
 * link, import or execute this.  Allowed users: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs (as the checker or the
 * reported CPU baseline, never as the thing shipped).
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose arithmetic it restates.  Dot products and reductions accumulate in
 * double: the reference accumulates the same products in F32 in a lane-strided
 * order (src/cuda/gemm.cu), so this restatement sits inside F32 rounding of any
 * summation order and can arbitrate between the reference CUDA build and ours.
 *
 * Parity pinning: checked against the reference's own golden vectors
 * (tests/test_gemm.cpp, tests/test_tensor.cpp), against golden fixtures produced
 * by importing the reference's numpy dequantisers (tools/decompose_gguf.py:219-378;
 * tests/golden/, generator tests/golden/make_golden.py) and, on the GPU box,
 * against the reference's CUDA kernels compiled from source (oracle/_ref).
 */
#ifndef NT_ORACLE_H
#define NT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DType ids — src/core/types.h:24-35 */
enum {
    NTO_F32 = 0, NTO_F16 = 1, NTO_Q8_0 = 2, NTO_Q4_0 = 3,
    NTO_Q4_K = 4, NTO_Q6_K = 5, NTO_Q5_K = 6
};

/* src/core/types.h:38-88 */
size_t nto_dtype_size(int dtype);
size_t nto_dtype_block_size(int dtype);
size_t nto_row_bytes(int dtype, int64_t n);

/* src/model/transformer.cpp:394-417 (fp16_to_fp32) and the RN conversion the
 * reference gets from __float2half (src/cuda/attention.cu:338-339). */
float    nto_fp16_to_fp32(uint16_t h);
uint16_t nto_fp32_to_fp16(float f);

/* Dequantise one row of n weights. Q8_0/Q4_0/Q6_K/Q4_K follow
 * src/model/transformer.cpp:449-594; Q5_K follows src/cuda/gemm.cu:300-350 and
 * tools/decompose_gguf.py:318-367; F16/F32 transformer.cpp:428-448. */
void nto_dequant_row(int dtype, const void* row, int64_t n, float* out);

/* y[out] = W[out,in] . x  — src/cuda/gemm.cu:32-671 (per-format kernels),
 * launcher gemm.cu:748-805.  Same per-block formulas, double accumulation. */
void nto_gemv(float* y, const void* W, const float* x, int out_features,
              int in_features, int dtype);

/* src/cuda/rmsnorm.cu:17-70 */
void nto_rmsnorm(float* y, const float* x, const float* w, int rows,
                 int hidden, float eps);

/* src/cuda/rotary.cu:16-62 (half-split pairs) / 65-107 (interleaved) */
void nto_rope(float* q, float* k, const int* positions, int seq_len,
              int n_heads, int n_kv_heads, int head_dim, float theta_base,
              float freq_scale, int interleaved);

/* src/cuda/attention.cu:316-342 */
void nto_copy_to_kv_cache(uint16_t* k_cache, uint16_t* v_cache, const float* k,
                          const float* v, int seq_len, int n_kv_heads,
                          int head_dim, int start_pos, int max_seq);

/* src/cuda/attention.cu:108-202 */
void nto_attention_decode(float* out, const float* q, const uint16_t* k_cache,
                          const uint16_t* v_cache, int seq_len, int n_heads,
                          int n_kv_heads, int head_dim, int max_seq, float scale);

/* src/cuda/attention.cu:216-311 */
void nto_attention_prefill(float* out, const float* Q, const uint16_t* k_cache,
                           const uint16_t* v_cache, int seq_len, int start_pos,
                           int n_heads, int n_kv_heads, int head_dim,
                           int max_seq, float scale);

/* src/cuda/gemm.cu:713-725 ; src/cuda/elementwise.cu:23-32 */
void nto_silu_mul(float* out, const float* gate, const float* up, int n);
void nto_add_inplace(float* a, const float* b, int n);

/* ---- whole-model forward: src/model/transformer.cpp:604-669 ---- */
typedef struct {
    int vocab_size, hidden_size, intermediate_size, n_layers;
    int n_heads, n_kv_heads, head_dim, max_seq_len;
    float norm_eps, rope_theta;
} nto_config;

typedef struct {
    const float* attn_norm;  const float* ffn_norm;
    const void *wq, *wk, *wv, *wo, *w_gate, *w_up, *w_down;
    int dt_q, dt_k, dt_v, dt_o, dt_gate, dt_up, dt_down;
} nto_layer;

typedef struct nto_model nto_model;

nto_model* nto_model_create(const nto_config* cfg);
void       nto_model_destroy(nto_model* m);
/* Pointers are borrowed (caller keeps the buffers alive). */
void nto_model_set_globals(nto_model* m, const void* token_embd, int dt_embd,
                           const void* output_w, int dt_output,
                           const float* output_norm);
void nto_model_set_layer(nto_model* m, int i, const nto_layer* l);
/* Runs seq_len tokens starting at start_pos; writes vocab logits of the last
 * token.  n_layers_run < 0 means all layers; 0..n runs that many (bench's bounded
 * cpu sample runs fewer). Returns 0 on success. */
int nto_model_forward(nto_model* m, const int* tokens, int seq_len,
                      int start_pos, float* logits, int n_layers_run);
int nto_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
