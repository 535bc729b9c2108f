// elementwise.cu — activation quantiser, RMSNorm, RoPE, KV-cache write, SwiGLU, residual adds, softmax,
// embedding gather.  Reference kernels K10-K15, K19-K21 (src/cuda/{rmsnorm,rotary,elementwise,softmax}.cu,
// gemm.cu:699-725, attention.cu:316-342).
//
// This file is compiled with --use_fast_math like the reference (CMakeLists.txt:20) and writes the
// transcendental expressions in the same form (1.0f / powf, cosf/sinf, g / (1 + expf(-g)), rsqrtf) so they
// lower to the same approximate instructions: greedy-token parity depends on it.
// Anything that must stay IEEE uses explicit *_rn intrinsics.
#include "kernels_internal.h"
#include "xquant.cuh"
#include "ring.cuh"
#include <cuda_fp16.h>
#include <cfloat>
#include <atomic>

namespace nt { namespace b200 {

static std::atomic<unsigned long long> g_launches{0};
unsigned long long launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }
static std::atomic<bool> g_pdl{false};
void set_pdl(bool on) { g_pdl.store(on); }
bool pdl_enabled() { return g_pdl.load(); }

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
// Block-wide sum / max over blockDim.x (multiple of 32, <= 1024) threads; result valid in every thread.
__device__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.f;
    return warp_sum(t);
}
__device__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : -FLT_MAX;
    return warp_max(t);
}

__global__ void quantize_x_kernel(const float* __restrict__ x, int8_t* __restrict__ xq, int K) {
    pdl_launch_dependents();
    pdl_wait();
    const int blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (blk * 32 >= K) return;
    quantize_block32(x[blk * 32 + lane], blk, lane, xq, K);
}

// ---- RMSNorm: y = x * rsqrtf(mean(x^2) + eps) * w, one CTA per row (rmsnorm.cu:17-70) ----
template <bool HALF_OUT>
__global__ void __launch_bounds__(1024) rmsnorm_kernel(void* __restrict__ yv, const float* __restrict__ xin,
                                                       const float* __restrict__ w, int hidden, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[32];
    const float* x = xin + (size_t)blockIdx.x * hidden;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) { float v = x[i]; ss += v * v; }
    ss = block_sum(ss, red);
    float mean_sq = ss / hidden;
    float rms_inv = rsqrtf(mean_sq + eps);
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
        const float v = x[i] * rms_inv * w[i];
        if (HALF_OUT) reinterpret_cast<__half*>(yv)[(size_t)blockIdx.x * hidden + i] = __float2half(v);
        else reinterpret_cast<float*>(yv)[(size_t)blockIdx.x * hidden + i] = v;
    }
}

// Multi-CTA RMSNorm + activation quantiser for the decode step: every CTA recomputes the (cheap, L2-resident)
// sum of squares in the same order, so all CTAs derive the bit-identical rms_inv, then quantises 8 of the
// row's 32-element blocks.  ~3 us instead of the 13.7 us of the single-CTA form on hidden = 8192.
__global__ void __launch_bounds__(256) rmsnorm_xq_kernel(float* __restrict__ y, int8_t* __restrict__ xq, const float* __restrict__ x,
                                                         const float* __restrict__ w, int hidden, float eps) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += 256) { float v = x[i]; ss += v * v; }
    ss = block_sum(ss, red);
    const float rms_inv = rsqrtf(ss / hidden + eps);
    const int blk = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blk * 32 >= hidden) return;
    const int e = blk * 32 + lane;
    const float v = x[e] * rms_inv * w[e];
    if (y) y[e] = v;
    quantize_block32(v, blk, lane, xq, hidden);
}

// ---- RoPE (rotary.cu:16-107) ----
template <bool INTERLEAVED>
__global__ void rope_kernel(float* __restrict__ q, float* __restrict__ k, const int* __restrict__ positions, int seq_len,
                            int n_heads, int n_kv_heads, int head_dim, float theta_base, float freq_scale) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half_dim = head_dim / 2;
    const int total_q = seq_len * n_heads * half_dim, total_k = seq_len * n_kv_heads * half_dim;
    if (idx >= total_q + total_k) return;
    const bool is_key = idx >= total_q;
    const int li = is_key ? idx - total_q : idx;
    const int n_h = is_key ? n_kv_heads : n_heads;
    const int pair = li % half_dim, head = (li / half_dim) % n_h, sp = li / (half_dim * n_h);
    const int pos = positions[sp];
    float freq = 1.0f / powf(theta_base, (2.0f * pair) / head_dim);
    float angle = pos * freq * freq_scale;
    float c = cosf(angle), sn = sinf(angle);
    float* d = (is_key ? k : q) + (size_t)sp * n_h * head_dim + (size_t)head * head_dim;
    const int i0 = INTERLEAVED ? 2 * pair : pair, i1 = INTERLEAVED ? 2 * pair + 1 : pair + half_dim;
    float x0 = d[i0], x1 = d[i1];
    d[i0] = x0 * c - x1 * sn;
    d[i1] = x1 * c + x0 * sn;
}

// ---- KV-cache write: F32 -> F16 RN at [start_pos + s] (attention.cu:316-342) ----
__global__ void kv_write_kernel(__half* __restrict__ kc, __half* __restrict__ vc, const float* __restrict__ k,
                                const float* __restrict__ v, int seq_len, int row /* n_kv * hd */, int start_pos, int max_seq) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= seq_len * row) return;
    const int s = idx / row, c = idx - s * row;
    const int cp = start_pos + s;
    if (cp >= max_seq) return;                       // silently dropped, like attention.cu:336
    kc[(size_t)cp * row + c] = __float2half(k[idx]);
    vc[(size_t)cp * row + c] = __float2half(v[idx]);
}

// Decode-step fusion of K14 + K15: rotate q and k (pairs (i, i + hd/2)), then store F16 k, v at cache row *pos_dev.
__global__ void rope_kv_decode_kernel(float* __restrict__ q, float* __restrict__ k, const float* __restrict__ v,
                                      __half* __restrict__ kc, __half* __restrict__ vc, const int* __restrict__ pos_dev,
                                      int n_heads, int n_kv, int head_dim, float theta_base, float freq_scale, int max_seq) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half_dim = head_dim / 2;
    const int total_q = n_heads * half_dim, total_k = n_kv * half_dim;
    if (idx >= total_q + total_k) return;
    const bool is_key = idx >= total_q;
    const int li = is_key ? idx - total_q : idx;
    const int pair = li % half_dim, head = li / half_dim;
    const int pos = *pos_dev;
    float freq = 1.0f / powf(theta_base, (2.0f * pair) / head_dim);
    float angle = pos * freq * freq_scale;
    float c = cosf(angle), sn = sinf(angle);
    float* d = (is_key ? k : q) + (size_t)head * head_dim;
    float x0 = d[pair], x1 = d[pair + half_dim];
    float r0 = x0 * c - x1 * sn, r1 = x1 * c + x0 * sn;
    d[pair] = r0;
    d[pair + half_dim] = r1;
    if (is_key && pos < max_seq) {
        const size_t row = (size_t)pos * n_kv * head_dim + (size_t)head * head_dim;
        kc[row + pair] = __float2half(r0);
        kc[row + pair + half_dim] = __float2half(r1);
        vc[row + pair] = __float2half(v[(size_t)head * head_dim + pair]);
        vc[row + pair + half_dim] = __float2half(v[(size_t)head * head_dim + pair + half_dim]);
    }
}

__global__ void silu_mul_kernel(float* __restrict__ out, const float* __restrict__ gate, const float* __restrict__ up, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float g = gate[i]; float silu = g / (1.0f + expf(-g)); out[i] = silu * up[i]; }
}
__global__ void add_kernel(float* out, const float* a, const float* b, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
__global__ void copy_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
// single-block cosine similarity (elementwise.cu:46-84): dot / (|a||b|), 0 when the denominator is <= 1e-8
__global__ void __launch_bounds__(1024) cosine_kernel(float* result, const float* __restrict__ a, const float* __restrict__ b, int n) {
    __shared__ float red[32];
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { float x = a[i], y = b[i]; dot += x * y; na += x * x; nb += y * y; }
    dot = block_sum(dot, red); na = block_sum(na, red); nb = block_sum(nb, red);
    if (threadIdx.x == 0) { float denom = sqrtf(na) * sqrtf(nb); *result = (denom > 1e-8f) ? dot / denom : 0.0f; }
}

// ---- row softmax, optional bool mask (softmax.cu:17-162): masked-out entries -> 0 ----
template <bool MASKED>
__global__ void __launch_bounds__(1024) softmax_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                       const bool* __restrict__ mask, int cols) {
    __shared__ float red[32];
    const float* x = in + (size_t)blockIdx.x * cols;
    const bool* mk = (MASKED && mask) ? mask + (size_t)blockIdx.x * cols : nullptr;
    float* y = out + (size_t)blockIdx.x * cols;
    float mx = -FLT_MAX;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) if (!mk || mk[i]) mx = fmaxf(mx, x[i]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
        float e = (!mk || mk[i]) ? expf(x[i] - mx) : 0.f;
        y[i] = e;
        sum += e;
    }
    sum = block_sum(sum, red);
    float inv = (sum > 0.f) ? 1.0f / sum : 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) y[i] *= inv;
}

// ---- C[M,N] = A[M,K] . B[N,K]^T in F32 (gemm.cu:677-694; dead code in the reference) ----
__global__ void gemm_f32_kernel(float* __restrict__ C, const float* __restrict__ A, const float* __restrict__ B, int M, int N, int K) {
    __shared__ float sa[16][17], sb[16][17];
    const int row = blockIdx.y * 16 + threadIdx.y, col = blockIdx.x * 16 + threadIdx.x;
    float acc = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
        int ka = k0 + threadIdx.x;
        sa[threadIdx.y][threadIdx.x] = (row < M && ka < K) ? A[(size_t)row * K + ka] : 0.f;
        int brow = blockIdx.x * 16 + threadIdx.y;
        sb[threadIdx.y][threadIdx.x] = (brow < N && ka < K) ? B[(size_t)brow * K + ka] : 0.f;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) acc = fmaf(sa[threadIdx.y][kk], sb[threadIdx.x][kk], acc);
        __syncthreads();
    }
    if (row < M && col < N) C[(size_t)row * N + col] = acc;
}

// ---- embedding gather + dequant on the GPU: one CTA per token ----
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ float dequant_at(const uint8_t* row, int dt, int i) {
    switch (dt) {
        case (int)DType::F32: return reinterpret_cast<const float*>(row)[i];
        case (int)DType::F16: return h2f(reinterpret_cast<const uint16_t*>(row)[i]);
        case (int)DType::Q8_0: {
            const uint8_t* b = row + (size_t)(i >> 5) * 34;
            return h2f(*reinterpret_cast<const uint16_t*>(b)) * (float)(int8_t)b[2 + (i & 31)];
        }
        case (int)DType::Q4_0: {
            const uint8_t* b = row + (size_t)(i >> 5) * 18;
            int j = i & 31;
            uint8_t byte = b[2 + (j & 15)];
            int qv = (j < 16) ? (byte & 15) : (byte >> 4);
            return h2f(*reinterpret_cast<const uint16_t*>(b)) * (float)(qv - 8);
        }
        case (int)DType::Q4_K_M: {
            const uint8_t* b = row + (size_t)(i >> 8) * 144;
            int n = i & 255, chunk = n >> 6, l = n & 31, hi = (n >> 5) & 1;
            int is = 2 * chunk + hi;
            const uint8_t* s = b + 4;
            int sc, m;
            if (is < 4) { sc = s[is] & 63; m = s[is + 4] & 63; }
            else { sc = (s[is + 4] & 15) | ((s[is - 4] >> 6) << 4); m = (s[is + 4] >> 4) | ((s[is] >> 6) << 4); }
            uint8_t byte = b[16 + chunk * 32 + l];
            int qv = hi ? (byte >> 4) : (byte & 15);
            float d = h2f(*reinterpret_cast<const uint16_t*>(b)), dmin = h2f(*reinterpret_cast<const uint16_t*>(b + 2));
            return (d * sc) * qv - dmin * m;
        }
        case (int)DType::Q6_K: {
            const uint8_t* b = row + (size_t)(i >> 8) * 210;
            int n = i & 255, hf = n >> 7, r = n & 127, run = r >> 5, l = r & 31;
            const uint8_t* ql = b + 64 * hf;
            const uint8_t* qh = b + 128 + 32 * hf;
            const int8_t* sc = reinterpret_cast<const int8_t*>(b + 192 + 8 * hf);
            uint8_t qb = ql[l + ((run & 1) ? 32 : 0)];
            int lo = (run >= 2) ? (qb >> 4) : (qb & 15);
            int q = (lo | (((qh[l] >> (2 * run)) & 3) << 4)) - 32;
            float d = h2f(*reinterpret_cast<const uint16_t*>(b + 208));
            return d * (float)sc[(l >> 4) + 2 * run] * q;
        }
        default: return 0.f;        // Q5_K table: zeros, as the reference (transformer.cpp:595-598)
    }
}
__global__ void embed_kernel(float* __restrict__ out, const uint8_t* __restrict__ table, int dt, size_t row_bytes,
                             const int* __restrict__ tokens, int hidden) {
    pdl_launch_dependents();
    pdl_wait();
    const uint8_t* row = table + (size_t)tokens[blockIdx.x] * row_bytes;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) out[(size_t)blockIdx.x * hidden + i] = dequant_at(row, dt, i);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

void quantize_x(const float* x, void* xq, int K, cudaStream_t s) {
    NT_CHECK(K % 128 == 0, "quantize_x: K must be a multiple of 128");
    launch_k(quantize_x_kernel, dim3(cdiv(K, 256)), dim3(256), 0, s, x, static_cast<int8_t*>(xq), K);
    count_launch();
}

static int norm_threads(int hidden) { return hidden <= 1024 ? 256 : hidden <= 4096 ? 512 : 1024; }

void rmsnorm(float* y, const float* x, const float* w, int rows, int hidden, float eps, cudaStream_t s) {
    if (rows <= 0) return;
    launch_k(rmsnorm_kernel<false>, dim3(rows), dim3(norm_threads(hidden)), 0, s, (void*)y, x, w, hidden, eps);
    count_launch();
}
void rmsnorm_f16(void* y, const float* x, const float* w, int rows, int hidden, float eps, cudaStream_t s) {
    if (rows <= 0) return;
    rmsnorm_kernel<true><<<rows, norm_threads(hidden), 0, s>>>(y, x, w, hidden, eps);
    count_launch();
}
void rmsnorm_xq(float* y, void* xq, const float* x, const float* w, int hidden, float eps, cudaStream_t s) {
    NT_CHECK(hidden % 128 == 0, "rmsnorm_xq: hidden must be a multiple of 128");
    launch_k(rmsnorm_xq_kernel, dim3(cdiv(hidden / 32, 8)), dim3(256), 0, s, y, static_cast<int8_t*>(xq), x, w, hidden, eps);
    count_launch();
}
void rope(float* q, float* k, const int* positions, int seq_len, int n_heads, int n_kv_heads, int head_dim,
          float theta, float freq_scale, bool interleaved, cudaStream_t s) {
    int total = seq_len * (n_heads + n_kv_heads) * (head_dim / 2);
    if (total <= 0) return;
    if (interleaved) rope_kernel<true><<<cdiv(total, 256), 256, 0, s>>>(q, k, positions, seq_len, n_heads, n_kv_heads, head_dim, theta, freq_scale);
    else rope_kernel<false><<<cdiv(total, 256), 256, 0, s>>>(q, k, positions, seq_len, n_heads, n_kv_heads, head_dim, theta, freq_scale);
    count_launch();
}
void copy_to_kv_cache(void* kc, void* vc, const float* k, const float* v, int seq_len, int n_kv, int hd,
                      int start_pos, int max_seq, cudaStream_t s) {
    int total = seq_len * n_kv * hd;
    if (total <= 0) return;
    kv_write_kernel<<<cdiv(total, 256), 256, 0, s>>>(static_cast<__half*>(kc), static_cast<__half*>(vc), k, v, seq_len,
                                                     n_kv * hd, start_pos, max_seq);
    count_launch();
}
void rope_kv_decode(float* q, float* k, const float* v, void* kc, void* vc, const int* pos_dev, int n_heads, int n_kv,
                    int hd, float theta, float freq_scale, int max_seq, cudaStream_t s) {
    int total = (n_heads + n_kv) * (hd / 2);
    launch_k(rope_kv_decode_kernel, dim3(cdiv(total, 256)), dim3(256), 0, s, q, k, v, static_cast<__half*>(kc),
             static_cast<__half*>(vc), pos_dev, n_heads, n_kv, hd, theta, freq_scale, max_seq);
    count_launch();
}
void silu_mul(float* out, const float* gate, const float* up, int n, cudaStream_t s) {
    if (n <= 0) return;
    launch_k(silu_mul_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, out, gate, up, n);
    count_launch();
}
void add(float* out, const float* a, const float* b, int n, cudaStream_t s) {
    if (n <= 0) return;
    launch_k(add_kernel, dim3(cdiv(n, 256)), dim3(256), 0, s, out, a, b, n);
    count_launch();
}
void add_inplace(float* a, const float* b, int n, cudaStream_t s) { add(a, a, b, n, s); }
void add_bias(float* y, const float* b, int n, cudaStream_t s) { add(y, y, b, n, s); }
void copy(float* dst, const float* src, int n, cudaStream_t s) {
    if (n <= 0) return;
    copy_kernel<<<cdiv(n, 256), 256, 0, s>>>(dst, src, n);
    count_launch();
}
void cosine_similarity(float* result, const float* a, const float* b, int n, cudaStream_t s) {
    cosine_kernel<<<1, 1024, 0, s>>>(result, a, b, n);
    count_launch();
}
void softmax(float* out, const float* in, int rows, int cols, cudaStream_t s) {
    if (rows <= 0) return;
    softmax_kernel<false><<<rows, norm_threads(cols), 0, s>>>(out, in, nullptr, cols);
    count_launch();
}
void masked_softmax(float* out, const float* in, const bool* mask, int rows, int cols, cudaStream_t s) {
    if (rows <= 0) return;
    softmax_kernel<true><<<rows, norm_threads(cols), 0, s>>>(out, in, mask, cols);
    count_launch();
}
void gemm_f32(float* C, const float* A, const float* B, int M, int N, int K, cudaStream_t s) {
    if (M <= 0 || N <= 0) return;
    dim3 block(16, 16), grid(cdiv(N, 16), cdiv(M, 16));
    gemm_f32_kernel<<<grid, block, 0, s>>>(C, A, B, M, N, K);
    count_launch();
}
void embed_rows(float* out, const void* table, DType dt, const int* tokens_dev, int n_tokens, int hidden, cudaStream_t s) {
    if (n_tokens <= 0) return;
    launch_k(embed_kernel, dim3(n_tokens), dim3(256), 0, s, out, static_cast<const uint8_t*>(table), (int)dt,
             dtype_row_size(dt, (size_t)hidden), tokens_dev, hidden);
    count_launch();
}

}}  // namespace nt::b200
