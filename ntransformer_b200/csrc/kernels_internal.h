// kernels_internal.h — host-callable launch functions behind the C-ABI (include/nt_b200.h)
// and behind the nt::cuda::launch_* drop-in symbols. Everything here is sm_100a CUDA; there is
// deliberately no CPU path.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "nt_types.h"

namespace nt { namespace b200 {

// ---------------------------------------------------------------------------------------------
// Block-scaled 3-term int8 activations ("xq").  x[e] ~= scale[e/32] * (q1*16384 + q2*128 + q3)
// with q1 in [-127,127], q2,q3 in [-64,64]: |error| <= 2^-22 of the 32-block absmax, i.e. F32-class.
// Integer dot products against the 4/5/6/8-bit weight codes are exact (IDP.4A), the float work is
// one scale per 16/32 weights instead of one convert + FMA per weight.
//   layout for K elements (K % 128 == 0): [q1: K][q2: K][q3: K][scale: K/32 f32][sum16: K/16 f32]; inside each plane
//   element e lives at byte e ^ (((e >> 7) & 7) << 4)  (the GEMV's bank-conflict-free shared-memory order)
// sum16 holds exact F32 sums of x over 16-element groups (for the dmin / -32 offset terms).
// ---------------------------------------------------------------------------------------------
inline size_t xq_bytes(int K) { return (size_t)3 * K + (size_t)(K / 32) * 4 + (size_t)(K / 16) * 4; }
void quantize_x(const float* x, void* xq, int K, cudaStream_t s);

// One weight matrix inside a (possibly fused) GEMV launch.
struct GemvMat {
    const void* W = nullptr;   // row-major GGUF blocks
    float* y = nullptr;        // output rows
    int out = 0;               // rows
    DType dtype = DType::F32;
    size_t row_pitch = 0;      // bytes between rows (0 = dense GGUF rows)
};
enum GemvEpilogue { GEMV_STORE = 0, GEMV_ADD = 1 /* y += W.x (residual) */, GEMV_SWIGLU = 2 /* y0 = silu(W0.x) * (W1.x) */,
                    GEMV_PEER = 3 /* tensor parallel: rows go to every rank's slot over NVLink (PeerOut), y is not written */ };


// Fused K-quant GEMV over up to 3 matrices sharing one activation vector (already in xq form).
// All matrices must be Q4_K / Q5_K / Q6_K with 16-byte aligned W and row pitch.
// Where the shared activation vector comes from: either pre-quantised `xq`, or an F32 vector `x` that the
// kernel quantises in its prologue (staged over the not-yet-primed part of its TMA ring), optionally as
// RMSNorm(x) * norm_w: the prologue quantises x * norm_w, sums x^2 in the same pass and the scalar
// rsqrt(mean(x^2) + eps) is applied to the results.
// Where a tensor-parallel GEMV (o-projection / down-projection shard) delivers its partial rows (engine/peer_xchg.h): the slot
// [parity][rank][hidden] of every rank.  A slot element is ONE 8-byte word {sequence number : value bits} written with a single
// 64-bit store (single-copy atomic), so the data is its own arrival flag: no fence, no separate flag hop — the receiver polls
// the element until it carries the sequence number it expects.  Pointers index ranks; entry `rank` is this GPU's own buffer.
struct PeerOut {
    static constexpr int kMaxTP = 8;
    unsigned long long* slots[kMaxTP];
    unsigned* seq;                 // exchanges completed on this rank; the exchange in flight carries *seq + 1
    unsigned* abort_word;
    int rank, size, hidden;
};
struct GemvInput {
    const void* xq = nullptr;
    const float* x = nullptr;
    const float* norm_w = nullptr;
    float eps = 0.f;
    const PeerOut* peer = nullptr;     // epilogue GEMV_PEER only
};
bool gemv_kq_supported(const GemvMat* mats, int n_mat, int K);
void gemv_kq(const GemvMat* mats, int n_mat, int K, const GemvInput& in, GemvEpilogue ep, cudaStream_t s);
void gemv_kq(const GemvMat* mats, int n_mat, int K, const void* xq, GemvEpilogue ep, cudaStream_t s);
// gemv_kquant_q.cu: the quarter-block kernel (lane <-> 64 weights, up to 16 warps); false = shape not covered, nothing launched
bool gemv_kq_quarter(const GemvMat* mats, int n_mat, int K, const GemvInput& in, GemvEpilogue ep, const PeerOut* peer, cudaStream_t s);
// epilogue GEMV_PEER: one matrix, rows == peer.hidden
void gemv_kq_peer(const GemvMat& mat, int K, const GemvInput& in, const PeerOut& peer, cudaStream_t s);

// Generic GEMV for every dtype / any alignment, F32 activations (Q8_0, Q4_0, F16, F32 and
// unaligned K-quant shards).
void gemv_generic(float* y, const void* W, const float* x, int out, int in, DType dt, size_t row_pitch,
                  GemvEpilogue ep, cudaStream_t s);


// ---- elementwise / norm / rope / kv / attention (reference kernels K10-K21) ----
void rmsnorm(float* y, const float* x, const float* w, int rows, int hidden, float eps, cudaStream_t s);
void rmsnorm_f16(void* y, const float* x, const float* w, int rows, int hidden, float eps, cudaStream_t s);
// fused: y = rmsnorm(x) (optional, may be null) and xq = quantise(rmsnorm(x)); rows == 1
void rmsnorm_xq(float* y, void* xq, const float* x, const float* w, int hidden, float eps, cudaStream_t s);
void rope(float* q, float* k, const int* positions, int seq_len, int n_heads, int n_kv_heads, int head_dim,
          float theta, float freq_scale, bool interleaved, cudaStream_t s);
void copy_to_kv_cache(void* kc, void* vc, const float* k, const float* v, int seq_len, int n_kv, int hd,
                      int start_pos, int max_seq, cudaStream_t s);
void silu_mul(float* out, const float* gate, const float* up, int n, cudaStream_t s);
void add_bias(float* y, const float* b, int n, cudaStream_t s);
void add(float* out, const float* a, const float* b, int n, cudaStream_t s);
void add_inplace(float* a, const float* b, int n, cudaStream_t s);
void copy(float* dst, const float* src, int n, cudaStream_t s);
void cosine_similarity(float* result, const float* a, const float* b, int n, cudaStream_t s);
void softmax(float* out, const float* in, int rows, int cols, cudaStream_t s);
void masked_softmax(float* out, const float* in, const bool* mask, int rows, int cols, cudaStream_t s);
void gemm_f32(float* C, const float* A, const float* B, int M, int N, int K, cudaStream_t s);
void attention_decode(float* out, const float* q, const void* kc, const void* vc, int seq_len, int n_heads,
                      int n_kv, int hd, int max_seq, float scale, cudaStream_t s);
void attention_prefill(float* out, const float* Q, const void* kc, const void* vc, int seq_len, int start_pos,
                       int n_heads, int n_kv, int hd, int max_seq, float scale, cudaStream_t s);
// attention_prefill_mma.cu: 64-query x 64-key tiles on mma.sync (used by attention_prefill for seq_len >= 16, hd 64/128)
bool attention_prefill_mma_supported(int seq_len, int n_heads, int n_kv, int hd);
void attention_prefill_mma(float* out, const float* Q, const void* kc, const void* vc, int seq_len, int start_pos, int n_heads,
                           int n_kv, int hd, int max_seq, float scale, cudaStream_t s);
// CUDA-graph friendly decode attention: context length = *pos_dev + 1 is read on the device.
int attention_decode_dyn_splits(int max_seq, int n_heads, int n_kv);
size_t attention_decode_dyn_scratch_floats(int max_seq, int n_heads, int n_kv, int hd);
// xq_out (optional): the merged output is also written in xq form for the o-projection GEMV.
void attention_decode_dyn(float* out, const float* q, const void* kc, const void* vc, const int* pos_dev, int max_seq,
                          int n_heads, int n_kv, int hd, float scale, float* scratch, void* xq_out, cudaStream_t s);
// The decode step's attention sub-block in one launch: RoPE (q and the new key; q, k are NOT modified in memory) + F16 KV-cache
// write at row *pos_dev + split-context attention + merge of the splits + optional xq of the result.  scratch as for
// attention_decode_dyn; tickets: attention_decode_fused_tickets() zeroed words (left zeroed).  hd 64 / 128 / 256.
int attention_decode_fused_tickets(int n_heads, int n_kv);
void attention_decode_fused(float* out, const float* q, const float* k, const float* v, void* kc, void* vc, const int* pos_dev,
                            int max_seq, int n_heads, int n_kv, int hd, float theta, float freq_scale, float scale, float* scratch,
                            unsigned* tickets, void* xq_out, cudaStream_t s);
// Two-launch form: RoPE + KV-cache write inside the split-context decode kernel (q, k not modified in memory), then the merge kernel.
void attention_decode_rope_dyn(float* out, const float* q, const float* k, const float* v, void* kc, void* vc, const int* pos_dev,
                               int max_seq, int n_heads, int n_kv, int hd, float theta, float freq_scale, float scale, float* scratch,
                               void* xq_out, cudaStream_t s);
// Fused RoPE (q, k in place; reference rotary.cu:16-62, non-interleaved pairs) + F16 KV-cache write at *pos_dev.
void rope_kv_decode(float* q, float* k, const float* v, void* kc, void* vc, const int* pos_dev, int n_heads, int n_kv,
                    int hd, float theta, float freq_scale, int max_seq, cudaStream_t s);
// Embedding-row gather + dequant on the GPU (reference does this on the CPU, transformer.cpp:419-599).
void embed_rows(float* out, const void* table, DType dt, const int* tokens_dev, int n_tokens, int hidden, cudaStream_t s);

// sample.cu: repeat penalty (in place on the logits) + temperature + top-k + top-p + inverse-CDF draw with the host-supplied
// uniform variate r; the sampled token id goes to out_dev[0].  Returns false (nothing launched) unless 0 < top_k <= 1024 < n
// and temperature > 0 — the caller then samples on the host like the reference (sampler.cpp:47-117).
bool sample_topk_supported(int n, float temperature, int top_k);
bool sample_topk(float* logits_dev, int n, float temperature, int top_k, float top_p, float repeat_penalty, const int* recent_dev,
                 int n_recent, float r, int* out_dev, cudaStream_t s);

// Programmatic dependent launch (PDL): when enabled, kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that a kernel's prologue (barrier init, TMA weight
// prefetch) overlaps the tail of its predecessor; every kernel executes griddepcontrol.wait before it touches
// data produced upstream, so ordering is unchanged.  Off by default; the engine turns it on for decode graphs.
void set_pdl(bool on);
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    NT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// prefill_gemm.cu: tcgen05/TMEM GEMM, C[M,N] = A[M,K] (F32, split hi+lo F16) . W[N,K]^T (F16)
size_t gemm_f16_tc_workspace_bytes(int M, int K);
bool gemm_f16_tc(float* C, const float* A, const void* W_f16, int M, int N, int K, void* workspace, cudaStream_t s);
bool gemm_f16_tc_supported(const void* W_f16, int N, int K, size_t row_pitch);
void split_activations(void* workspace, const float* A, int M, int K, cudaStream_t s);
bool gemm_f16_tc_ws(float* C, const void* workspace, const void* W_f16, int M, int N, int K, bool add, cudaStream_t s);
// workspace_out <- split(silu(A.Wgate^T) * (A.Wup^T)): the FFN's gate/up GEMMs with the SwiGLU + split fused in the epilogue
bool gemm_f16_tc_swiglu_ws(void* workspace_out, const void* workspace_in, const void* Wgate_f16, const void* Wup_f16, int M, int N, int K,
                           cudaStream_t s);
void rmsnorm_split(void* workspace, const float* x, const float* w, int rows, int hidden, float eps, cudaStream_t s);
// prefill_dequant.cu: any GGUF weight matrix -> dense F16 pair with W = w_hi + w_lo (operands of two tensor-core GEMMs)
bool dequant_split_supported(DType dt);
void dequant_split(void* w_hi, void* w_lo, const void* W, DType dt, size_t row_pitch, int rows, int cols, cudaStream_t s);

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is per DEVICE, so the "already done" state is a bit per device
// id (a process that drives a second GPU must opt in there too).  `done` is the call site's static mask.
template <typename K>
inline void opt_in_dynamic_smem(K* kernel, int bytes, unsigned long long& done) {
    int dev = 0;
    NT_CUDA_CHECK(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (done & bit) return;
    NT_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done |= bit;
}
// After a raw <<<>>> launch: a rejected configuration must not pass silently.
#define NT_LAUNCH_CHECK() NT_CUDA_CHECK(cudaPeekAtLastError())

// Number of kernels launched by this library since load (bench.py's gpu_launches claim).
unsigned long long launch_count();
void count_launch(int n = 1);

}}  // namespace nt::b200
