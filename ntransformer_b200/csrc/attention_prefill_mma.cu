// attention_prefill_mma.cu — causal GQA attention over a whole prompt chunk (prefill), sm_100a.
//
// Same contract as the reference's attention_prefill_kernel (src/cuda/attention.cu:216-311): F32 queries
// [seq][n_heads][hd], F16 KV cache [max_seq][n_kv][hd], query i attends keys 0..start_pos+i, softmax(q.k*scale).v.
// The reference (and csrc/attention.cu's prefill_kernel) re-reads the whole key range once per query; here a CTA owns
// 64 queries of one head and streams 64-key K/V tiles through shared memory (cp.async, double buffered), with the two
// contractions on the warp-level tensor-core path (mma.sync m16n8k16, F32 accumulate) and an online softmax in
// registers (flash-attention-2 data flow).  F32 operands are split into two F16 terms (q = hi + lo, p = hi + lo) so
// the result keeps ~22 mantissa bits and stays within the 2e-5 tolerance of the oracle.
#include "kernels_internal.h"
#include <cuda_fp16.h>

namespace nt { namespace b200 {

namespace {

constexpr int BQ = 64, BKV = 64, PAD = 8;      // PAD halfs: row stride 272 B (hd 128) keeps LDS.32 / ldmatrix conflict-free

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// (x, y) -> packed F16 hi pair and lo pair with x = hi + lo to ~2^-22
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 f = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(x - f.x, y - f.y));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(128, 2) prefill_mma_kernel(float* __restrict__ out, const float* __restrict__ Q,
                                                             const __half* __restrict__ kc, const __half* __restrict__ vc,
                                                             int seq_len, int start_pos, int n_heads, int n_kv, int max_seq,
                                                             float scale) {
    constexpr int LD = HD + PAD, KS = HD / 16, NT = HD / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sK = reinterpret_cast<__half*>(smem_raw);                 // [2][BKV][LD]
    __half* sV = sK + 2 * BKV * LD;                                   // [2][BKV][LD]
    const int qb = gridDim.x - 1 - blockIdx.x;                        // heaviest (latest) query blocks first
    const int head = blockIdx.y, kvh = head / (n_heads / n_kv);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r0 = qb * BQ + warp * 16 + g, r1 = r0 + 8;              // this thread's two query rows (chunk-relative)
    const int p0 = start_pos + r0, p1 = start_pos + r1;               // their absolute positions == last visible key

    // ---- Q fragments (A operand, 16 x HD per warp), split into F16 hi + lo ----
    uint32_t qhi[KS][4], qlo[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = (i & 1) ? r1 : r0, c = ks * 16 + 2 * t + ((i & 2) ? 8 : 0);
            float2 v = make_float2(0.f, 0.f);
            if (r < seq_len) v = *reinterpret_cast<const float2*>(Q + ((size_t)r * n_heads + head) * HD + c);
            split2(v.x, v.y, qhi[ks][i], qlo[ks][i]);
        }
    }
    float o[NT][4];
#pragma unroll
    for (int n = 0; n < NT; n++) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    const int last_q = min(qb * BQ + BQ, seq_len) - 1;                // last valid query of the block
    const int n_tiles = (start_pos + last_q) / BKV + 1;               // key tiles covering positions 0..start_pos+last_q

    auto load_tile = [&](int j, int st) {
        constexpr int CH = HD / 8;                                    // 16-byte chunks per row
        for (int c = threadIdx.x; c < BKV * CH; c += 128) {
            const int row = c / CH, ch = c % CH;
            const int key = min(j * BKV + row, max_seq - 1);          // rows past the cache are masked by causality anyway
            const size_t src = ((size_t)key * n_kv + kvh) * HD + ch * 8;
            cp_async16(sK + ((size_t)st * BKV + row) * LD + ch * 8, kc + src);
            cp_async16(sV + ((size_t)st * BKV + row) * LD + ch * 8, vc + src);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    load_tile(0, 0);
    for (int j = 0; j < n_tiles; j++) {
        const int st = j & 1;
        if (j + 1 < n_tiles) {
            load_tile(j + 1, st ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const __half* tK = sK + (size_t)st * BKV * LD;
        const __half* tV = sV + (size_t)st * BKV * LD;

        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; n++) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
#pragma unroll
            for (int n = 0; n < 8; n++) {
                const __half* kp = tK + (size_t)(n * 8 + g) * LD + ks * 16 + 2 * t;
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kp), b1 = *reinterpret_cast<const uint32_t*>(kp + 8);
                mma16816(s[n], qhi[ks], b0, b1);
                mma16816(s[n], qlo[ks], b0, b1);
            }
        }
        // ---- scale, causal mask, online softmax ----
        const int key0 = j * BKV;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; n++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int key = key0 + n * 8 + 2 * t + (i & 1);
                const int p = (i & 2) ? p1 : p0;
                const float v = key <= p ? s[n][i] * scale : -INFINITY;
                s[n][i] = v;
                if (i & 2) mx1 = fmaxf(mx1, v); else mx0 = fmaxf(mx0, v);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);      // finite from tile 0 on: key 0 is visible to every row
        const float a0 = __expf(m0 - mn0), a1 = __expf(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int n = 0; n < 8; n++) {
            s[n][0] = __expf(s[n][0] - mn0); s[n][1] = __expf(s[n][1] - mn0);
            s[n][2] = __expf(s[n][2] - mn1); s[n][3] = __expf(s[n][3] - mn1);
            sum0 += s[n][0] + s[n][1]; sum1 += s[n][2] + s[n][3];
        }
        l0 = l0 * a0 + sum0; l1 = l1 * a1 + sum1;                    // per-thread partial row sums (reduced at the end)
#pragma unroll
        for (int n = 0; n < NT; n++) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }

        // ---- O += P V : the accumulator layout of S is the A-operand layout of the second MMA ----
#pragma unroll
        for (int kt = 0; kt < BKV / 16; kt++) {
            uint32_t phi[4], plo[4];
            split2(s[2 * kt][0], s[2 * kt][1], phi[0], plo[0]);
            split2(s[2 * kt][2], s[2 * kt][3], phi[1], plo[1]);
            split2(s[2 * kt + 1][0], s[2 * kt + 1][1], phi[2], plo[2]);
            split2(s[2 * kt + 1][2], s[2 * kt + 1][3], phi[3], plo[3]);
#pragma unroll
            for (int np = 0; np < HD / 16; np++) {
                // four 8x8 blocks of V^T: (keys +0..7 | +8..15) x (hd +0..7 | +8..15)
                const __half* vp = tV + (size_t)(kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + np * 16 + (lane >> 4) * 8;
                uint32_t v0, v1, v2, v3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"((uint32_t)__cvta_generic_to_shared(vp)));
                mma16816(o[2 * np], phi, v0, v1);
                mma16816(o[2 * np], plo, v0, v1);
                mma16816(o[2 * np + 1], phi, v2, v3);
                mma16816(o[2 * np + 1], plo, v2, v3);
            }
        }
        __syncthreads();                                             // this stage is overwritten by the prefetch of tile j + 2
    }
    l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 1); l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 2);
    l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 1); l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int n = 0; n < NT; n++) {
        const int c = n * 8 + 2 * t;
        if (r0 < seq_len) *reinterpret_cast<float2*>(out + ((size_t)r0 * n_heads + head) * HD + c) = make_float2(o[n][0] * i0, o[n][1] * i0);
        if (r1 < seq_len) *reinterpret_cast<float2*>(out + ((size_t)r1 * n_heads + head) * HD + c) = make_float2(o[n][2] * i1, o[n][3] * i1);
    }
}

template <int HD>
void launch(float* out, const float* Q, const __half* kc, const __half* vc, int seq_len, int start_pos, int n_heads, int n_kv,
            int max_seq, float scale, cudaStream_t s) {
    const int smem = 2 * 2 * BKV * (HD + PAD) * (int)sizeof(__half);
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(prefill_mma_kernel<HD>, (int)(smem), configured);
    prefill_mma_kernel<HD><<<dim3((seq_len + BQ - 1) / BQ, n_heads), 128, smem, s>>>(out, Q, kc, vc, seq_len, start_pos, n_heads, n_kv,
                                                                                    max_seq, scale);
    count_launch();
}

}  // namespace

bool attention_prefill_mma_supported(int seq_len, int n_heads, int n_kv, int hd) {
    return seq_len >= 16 && (hd == 64 || hd == 128) && n_kv > 0 && n_heads % n_kv == 0;
}

void attention_prefill_mma(float* out, const float* Q, const void* kc, const void* vc, int seq_len, int start_pos, int n_heads,
                           int n_kv, int hd, int max_seq, float scale, cudaStream_t s) {
    const __half* k = static_cast<const __half*>(kc);
    const __half* v = static_cast<const __half*>(vc);
    if (hd == 128) launch<128>(out, Q, k, v, seq_len, start_pos, n_heads, n_kv, max_seq, scale, s);
    else launch<64>(out, Q, k, v, seq_len, start_pos, n_heads, n_kv, max_seq, scale, s);
}

}}  // namespace nt::b200
