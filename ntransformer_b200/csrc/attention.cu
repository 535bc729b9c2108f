// attention.cu — GQA KV-cache attention for decode and prefill on sm_100a.
//
// Replaces reference kernels K16 (attention_decode_generic_kernel, src/cuda/attention.cu:108-202) and
// K18 (attention_prefill_kernel, :216-311).  Same math (F32 q, F16 K/V, scores * scale, softmax with the
// fast-math exp, P.V in F32) with a different work split:
//   * one CTA serves GC query heads that share a KV head, so K/V rows are fetched once per GQA group
//     instead of once per head (the reference re-reads them n_heads/n_kv_heads times);
//   * decode is split along the context (flash-decoding): grid = head-groups x splits so 148 SMs are
//     busy at ctx 2048 where the reference launches only n_heads CTAs; partial (max, sum, P.V) are merged
//     by a small combine kernel;
//   * K rows are read coalesced (one row per warp step, DPL dims per lane) and the GC partial dot
//     products are reduced with a transpose-reduce (GC-1 + log2(32/GC) shuffles instead of 5*GC).
// Compiled with --use_fast_math (expf -> ex2.approx path, as the reference build).
#include "kernels_internal.h"
#include "ring.cuh"
#include "xquant.cuh"
#include <cuda_fp16.h>
#include <algorithm>
#include <cfloat>
#include <mutex>

namespace nt { namespace b200 {

namespace {

constexpr int AW = 8;                    // warps per CTA
constexpr int MAX_SPLITS = 64;
// Graph-replayed decode (launch shape fixed by max_seq): a context slice holds at least this many keys.  Without the floor a short
// context is cut into one-key slices (ctx 64 over 64 slices when a rank holds a single KV head) and the merge walks them all:
// decode_combine took 13.9 us at tensor-parallel-8 shapes and 5.4 us at one GPU (profiles/r02_launches_tp8_shard.txt).
constexpr int DYN_MIN_SPLIT = 32;
constexpr int FUSED_MIN_SPLIT = 64;      // decode_fused_kernel: at least this many keys per context slice (one slice = no merge step)
constexpr int KU = 8;                    // cache rows a warp keeps in flight in the score and P.V loops
constexpr int ATTN_MAX_DYN_SMEM = 227 * 1024 - 1024;   // static __shared__ (s_max/s_sum) counts against the 227 KB cap

template <int N> struct HalfVec;         // N halfs loaded as one vector
template <> struct HalfVec<2> { using T = uint32_t; };
template <> struct HalfVec<4> { using T = uint2; };
template <> struct HalfVec<8> { using T = uint4; };

template <int DPL>
__device__ __forceinline__ void load_row(const __half* p, float (&f)[DPL]) {
    typename HalfVec<DPL>::T raw = __ldg(reinterpret_cast<const typename HalfVec<DPL>::T*>(p));
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < DPL / 2; i++) { float2 t = __half22float2(h2[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// Reduce GC per-lane partial sums across the warp. On return lane L holds (in v[0]) the full sum of
// value index (L / (32 / GC)) ... valid in every lane of that group.
template <int GC>
__device__ __forceinline__ float transpose_reduce(float (&v)[GC], int lane) {
    // exchange phase: halve the number of live values per step
#pragma unroll
    for (int n = GC, off = 16; n > 1; n >>= 1, off >>= 1) {
        const bool upper = lane & off;
#pragma unroll
        for (int i = 0; i < n / 2; i++) {
            float send = upper ? v[i] : v[i + n / 2];
            float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
        }
    }
    // plain butterfly over the remaining lanes of the group
    float r = v[0];
#pragma unroll
    for (int off = 16 / GC; off > 0; off >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, off);
    return r;
}

// Decode-step fusion (decode_fused_kernel): the token at cache row `pos` is not in the cache yet.  Its key arrives unrotated
// in F32 from the q/k/v GEMV; the warp that reaches row `pos` rotates it (reference rotary.cu:16-62, pairs (i, i + hd/2)), rounds
// key and value to F16 exactly as the cache write does (attention.cu:338-339) and — in the one CTA per KV head that owns the
// write — stores the row.  pos < 0: plain attention over the cache, queries used as given.
struct NewToken {
    const float* k = nullptr;      // [n_kv][HD]
    const float* v = nullptr;
    int pos = -1;
    float theta = 0.f, freq_scale = 1.f;
    __half* kc_w = nullptr;        // cache base pointers, non-null only in the CTA that stores the row
    __half* vc_w = nullptr;
};

// cos / sin of this lane's DPL rotation pairs.  Lane L holds dims [L*DPL, L*DPL + DPL): lanes 0-15 the first half of the
// head, lanes 16-31 the second, so a pair (i, i + hd/2) lives in lanes (L, L ^ 16) at the same register index.
// Same expression as rope_kv_decode_kernel / the reference (fast-math powf, cosf, sinf).
template <int DPL>
__device__ __forceinline__ void rope_angles(float (&c)[DPL], float (&sn)[DPL], int lane, int pos, float theta_base, float freq_scale) {
    constexpr int head_dim = DPL * 32;
#pragma unroll
    for (int i = 0; i < DPL; i++) {
        const int pair = (lane & 15) * DPL + i;
        float freq = 1.0f / powf(theta_base, (2.0f * pair) / head_dim);
        float angle = pos * freq * freq_scale;
        c[i] = cosf(angle); sn[i] = sinf(angle);
    }
}
template <int DPL>
__device__ __forceinline__ void rope_rotate(float (&x)[DPL], const float (&c)[DPL], const float (&sn)[DPL], int lane) {
#pragma unroll
    for (int i = 0; i < DPL; i++) {
        const float other = __shfl_xor_sync(0xFFFFFFFFu, x[i], 16);
        const bool lo = lane < 16;
        const float x0 = lo ? x[i] : other, x1 = lo ? other : x[i];
        const float r0 = x0 * c[i] - x1 * sn[i], r1 = x1 * c[i] + x0 * sn[i];
        x[i] = lo ? r0 : r1;
    }
}
template <int DPL>
__device__ __forceinline__ void store_row_f16(__half* p, const __half (&h)[DPL]) {
    typename HalfVec<DPL>::T raw;
    __half* d = reinterpret_cast<__half*>(&raw);
#pragma unroll
    for (int i = 0; i < DPL; i++) d[i] = h[i];
    *reinterpret_cast<typename HalfVec<DPL>::T*>(p) = raw;
}

// One CTA: GC query heads of one KV head, keys [k_begin, k_end).
// q_base: first of the GC heads' query vectors (contiguous [GC][hd]); out likewise.
// If part_* != nullptr writes unnormalised partials, else the normalised output.
template <int DPL, int GC>
__device__ void attend_group(float* __restrict__ out, const float* __restrict__ q_base, const __half* __restrict__ kc,
                             const __half* __restrict__ vc, int kv_head, int n_kv, int k_begin, int k_end, float scale,
                             float* __restrict__ part_o, float* __restrict__ part_ml, float* smem, const NewToken nt = NewToken{}) {
    constexpr int HD = DPL * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_keys = k_end - k_begin;
    float* sc = smem;                                   // [GC][n_keys]
    float* red = smem + (size_t)GC * n_keys;            // [AW][GC][HD]
    __shared__ float s_max[GC], s_sum[GC];
    const size_t row_stride = (size_t)n_kv * HD;
    const __half* kbase = kc + (size_t)kv_head * HD + (size_t)lane * DPL;
    const __half* vbase = vc + (size_t)kv_head * HD + (size_t)lane * DPL;

    float qr[GC][DPL];
#pragma unroll
    for (int g = 0; g < GC; g++)
#pragma unroll
        for (int i = 0; i < DPL; i++) qr[g][i] = q_base[(size_t)g * HD + lane * DPL + i];
    float rc[DPL], rs[DPL];
    if (nt.pos >= 0) {                                   // fused decode step: the queries arrive unrotated
        rope_angles<DPL>(rc, rs, lane, nt.pos, nt.theta, nt.freq_scale);
#pragma unroll
        for (int g = 0; g < GC; g++) rope_rotate<DPL>(qr[g], rc, rs, lane);
    }

    // ---- phase 1: scores (KU cache rows in flight per warp: one memory round trip per KU keys).  A slice of at most AW * KU
    // keys is one batch: its V rows are fetched in the same round trip and wait in registers for phase 3. ----
    const bool single = n_keys <= AW * KU;
    float vpre[KU][DPL];
    for (int p0 = warp; p0 < n_keys; p0 += AW * KU) {
        float kf[KU][DPL];
        if (single) {
#pragma unroll
            for (int u = 0; u < KU; u++) {
                const int p = p0 + u * AW;
                if (p >= n_keys) break;
                if (k_begin + p != nt.pos) load_row<DPL>(vbase + (size_t)(k_begin + p) * row_stride, vpre[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < KU; u++) {
            const int p = p0 + u * AW;
            if (p >= n_keys) break;                          // warp-uniform
            if (k_begin + p == nt.pos) {                     // warp-uniform: the new token's key, not in the cache yet
#pragma unroll
                for (int i = 0; i < DPL; i++) kf[u][i] = nt.k[(size_t)kv_head * HD + lane * DPL + i];
                rope_rotate<DPL>(kf[u], rc, rs, lane);
                __half kh[DPL];
#pragma unroll
                for (int i = 0; i < DPL; i++) { kh[i] = __float2half(kf[u][i]); kf[u][i] = __half2float(kh[i]); }
                if (nt.kc_w) store_row_f16<DPL>(nt.kc_w + (size_t)nt.pos * row_stride + (size_t)kv_head * HD + (size_t)lane * DPL, kh);
            } else {
                load_row<DPL>(kbase + (size_t)(k_begin + p) * row_stride, kf[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < KU; u++) {
            const int p = p0 + u * AW;
            if (p >= n_keys) break;
            float part[GC];
#pragma unroll
            for (int g = 0; g < GC; g++) {
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < DPL; i++) a = fmaf(qr[g][i], kf[u][i], a);
                part[g] = a;
            }
            float tot = transpose_reduce<GC>(part, lane);
            if ((lane & (32 / GC - 1)) == 0) sc[(size_t)(lane / (32 / GC)) * n_keys + p] = tot * scale;
        }
    }
    __syncthreads();
    // ---- phase 2: per-head max / exp / sum (warp g <-> head g) ----
    for (int g = warp; g < GC; g += AW) {
        float* s = sc + (size_t)g * n_keys;
        float mx = -FLT_MAX;
        for (int p = lane; p < n_keys; p += 32) mx = fmaxf(mx, s[p]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float sum = 0.f;
        for (int p = lane; p < n_keys; p += 32) { float e = expf(s[p] - mx); s[p] = e; sum += e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if (lane == 0) { s_max[g] = mx; s_sum[g] = sum; }
    }
    __syncthreads();
    // ---- phase 3: P.V (warp <-> keys, lane <-> DPL dims) ----
    float acc[GC][DPL];
#pragma unroll
    for (int g = 0; g < GC; g++)
#pragma unroll
        for (int i = 0; i < DPL; i++) acc[g][i] = 0.f;
    for (int p0 = warp; p0 < n_keys; p0 += AW * KU) {
        float vf[KU][DPL];
#pragma unroll
        for (int u = 0; u < KU; u++) {
            const int p = p0 + u * AW;
            if (p >= n_keys) break;
            if (k_begin + p == nt.pos) {                     // the new token's value: F16-rounded like the cache row it becomes
                __half vh[DPL];
#pragma unroll
                for (int i = 0; i < DPL; i++) { vh[i] = __float2half(nt.v[(size_t)kv_head * HD + lane * DPL + i]); vf[u][i] = __half2float(vh[i]); }
                if (nt.vc_w) store_row_f16<DPL>(nt.vc_w + (size_t)nt.pos * row_stride + (size_t)kv_head * HD + (size_t)lane * DPL, vh);
            } else if (single) {
#pragma unroll
                for (int i = 0; i < DPL; i++) vf[u][i] = vpre[u][i];
            } else {
                load_row<DPL>(vbase + (size_t)(k_begin + p) * row_stride, vf[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < KU; u++) {
            const int p = p0 + u * AW;
            if (p >= n_keys) break;
#pragma unroll
            for (int g = 0; g < GC; g++) {
                float w = sc[(size_t)g * n_keys + p];
#pragma unroll
                for (int i = 0; i < DPL; i++) acc[g][i] = fmaf(w, vf[u][i], acc[g][i]);
            }
        }
    }
#pragma unroll
    for (int g = 0; g < GC; g++)
#pragma unroll
        for (int i = 0; i < DPL; i++) red[((size_t)warp * GC + g) * HD + lane * DPL + i] = acc[g][i];
    __syncthreads();
    for (int idx = threadIdx.x; idx < GC * HD; idx += AW * 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < AW; w++) t += red[(size_t)w * GC * HD + idx];
        const int g = idx / HD, d = idx - g * HD;
        if (part_o) {
            part_o[(size_t)g * HD + d] = t;             // caller strides heads by HD inside its split slot
        } else {
            float l = s_sum[g];
            out[(size_t)g * HD + d] = t * ((l > 0.f) ? 1.0f / l : 0.f);
        }
    }
    if (part_ml && threadIdx.x < GC) {
        part_ml[2 * threadIdx.x] = s_max[threadIdx.x];
        part_ml[2 * threadIdx.x + 1] = s_sum[threadIdx.x];
    }
}

// Scratch layout for split decode: [head][split][HD] floats then [head][split][2] (max, sum).
template <int DPL, int GC>
__global__ void __launch_bounds__(AW * 32) decode_kernel(float* __restrict__ out, const float* __restrict__ q,
                                                         const __half* __restrict__ kc, const __half* __restrict__ vc,
                                                         int seq_len, int n_heads, int n_kv, float scale, int n_splits,
                                                         int split_len, float* __restrict__ scratch,
                                                         const int* __restrict__ pos_dev) {
    constexpr int HD = DPL * 32;
    extern __shared__ float smem_dyn[];
    pdl_launch_dependents();
    pdl_wait();
    if (pos_dev) {                       // CUDA-graph replay: context length lives in device memory
        seq_len = *pos_dev + 1;
        split_len = max((seq_len + n_splits - 1) / n_splits, DYN_MIN_SPLIT);
    }
    const int head0 = blockIdx.x * GC, split = blockIdx.y;
    const int kv_head = head0 / (n_heads / n_kv);
    const int k_begin = split * split_len, k_end = min(seq_len, k_begin + split_len);
    if (k_begin >= k_end) return;
    if (n_splits == 1 && !pos_dev) {
        attend_group<DPL, GC>(out + (size_t)head0 * HD, q + (size_t)head0 * HD, kc, vc, kv_head, n_kv, k_begin, k_end, scale,
                              nullptr, nullptr, smem_dyn);
    } else {
        // partials for head (head0+g) live at scratch[((head0+g) * n_splits + split) * HD]; attend_group strides
        // heads by HD, so hand it a staging area in shared memory and scatter afterwards.
        float* stage = smem_dyn + (size_t)GC * (k_end - k_begin) + (size_t)AW * GC * HD;   // [GC][HD] + [GC][2]
        attend_group<DPL, GC>(nullptr, q + (size_t)head0 * HD, kc, vc, kv_head, n_kv, k_begin, k_end, scale, stage,
                              stage + GC * HD, smem_dyn);
        __syncthreads();
        float* ml = scratch + (size_t)n_heads * n_splits * HD;
        for (int idx = threadIdx.x; idx < GC * HD; idx += AW * 32) {
            const int g = idx / HD, d = idx - g * HD;
            scratch[((size_t)(head0 + g) * n_splits + split) * HD + d] = stage[idx];
        }
        if (threadIdx.x < 2 * GC) {
            const int g = threadIdx.x >> 1;
            ml[((size_t)(head0 + g) * n_splits + split) * 2 + (threadIdx.x & 1)] = stage[GC * HD + threadIdx.x];
        }
    }
}

// Merge split partials: out[h] = sum_i e^{m_i - m} o_i / sum_i e^{m_i - m} l_i
// xq_out (optional): also emits the block-scaled int8x3 form of the output vector (kernels_internal.h "xq") that
// the o-projection GEMV consumes, saving a separate quantise launch.  Requires blockDim.x == 128, hd % 32 == 0.
__global__ void decode_combine_kernel(float* __restrict__ out, const float* __restrict__ scratch, int n_heads, int hd,
                                      int n_splits, int seq_len, int split_len, const int* __restrict__ pos_dev,
                                      int8_t* __restrict__ xq_out) {
    __shared__ float w_s[MAX_SPLITS];                // e^{m_i - m} per slice
    __shared__ float inv_s;
    const int h = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (pos_dev) { seq_len = *pos_dev + 1; split_len = max((seq_len + n_splits - 1) / n_splits, DYN_MIN_SPLIT); }
    const float* ml = scratch + (size_t)n_heads * n_splits * hd + (size_t)h * n_splits * 2;
    const int used = (seq_len + split_len - 1) / split_len;
    // the slices' (max, sum) pairs: one lane per slice (two when there are more than 32), one memory round trip, warp reductions
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int i0 = lane, i1 = lane + 32;
        const float m0 = (i0 < used) ? ml[2 * i0] : -FLT_MAX, l0 = (i0 < used) ? ml[2 * i0 + 1] : 0.f;
        const float m1 = (i1 < used) ? ml[2 * i1] : -FLT_MAX, l1 = (i1 < used) ? ml[2 * i1 + 1] : 0.f;
        float m = fmaxf(m0, m1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        const float w0 = (i0 < used) ? expf(m0 - m) : 0.f, w1 = (i1 < used) ? expf(m1 - m) : 0.f;
        float l = l0 * w0 + l1 * w1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xFFFFFFFFu, l, o);
        if (i0 < MAX_SPLITS) w_s[i0] = w0;
        if (i1 < MAX_SPLITS) w_s[i1] = w1;
        if (lane == 0) inv_s = (l > 0.f) ? 1.0f / l : 0.f;
    }
    __syncthreads();
    const float inv = inv_s;
    for (int d = threadIdx.x; d < hd; d += blockDim.x) {       // hd % 32 == 0: whole warps stay together
        const float* sp = scratch + (size_t)h * n_splits * hd + d;
        float o = 0.f;
        int i = 0;
        for (; i + 8 <= used; i += 8) {                        // 8 independent loads in flight
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; j++) t[j] = sp[(size_t)(i + j) * hd];
#pragma unroll
            for (int j = 0; j < 8; j++) o = fmaf(t[j], w_s[i + j], o);
        }
        for (; i < used; i++) o = fmaf(sp[(size_t)i * hd], w_s[i], o);
        const float v = o * inv;
        out[(size_t)h * hd + d] = v;
        if (xq_out) {
            const int K = n_heads * hd, e = h * hd + d, lane = threadIdx.x & 31;
            int q1, q2, q3;
            float sc, s16;
            quantize_lane32(v, q1, q2, q3, sc, s16);
            const int se = (int)xq_swizzle((uint32_t)e);
            xq_out[se] = (int8_t)q1;
            xq_out[K + se] = (int8_t)q2;
            xq_out[2 * K + se] = (int8_t)q3;
            float* scale = reinterpret_cast<float*>(xq_out + 3 * (size_t)K);
            if (lane == 0) scale[e >> 5] = sc;
            if ((lane & 15) == 0) scale[K / 32 + (e >> 4)] = s16;
        }
    }
}

// The decode step's whole attention sub-block in ONE launch: RoPE of the queries and of the new key, the F16 KV-cache write
// at row *pos_dev, split-context GQA attention, and — by the CTA that finishes a head group's last split ("last arriver", one
// ticket per head group) — the merge of the splits and the xq form of the result for the o-projection.  Replaces
// rope_kv_decode_kernel + decode_kernel + decode_combine_kernel (reference K14 rotary.cu:16-62, K15 attention.cu:316-342,
// K16 attention.cu:108-202); the merge keeps decode_combine_kernel's arithmetic and order, so both paths agree bit for bit
// whenever they cut the context at the same places.
template <int DPL, int GC>
__global__ void __launch_bounds__(AW * 32) decode_fused_kernel(float* __restrict__ out, const float* __restrict__ q,
                                                               const float* __restrict__ k_new, const float* __restrict__ v_new,
                                                               __half* __restrict__ kc, __half* __restrict__ vc,
                                                               const int* __restrict__ pos_dev, int n_heads, int n_kv, float scale,
                                                               float theta, float freq_scale, int n_splits, int min_split,
                                                               float* __restrict__ scratch, unsigned* __restrict__ tickets,
                                                               int8_t* __restrict__ xq_out) {
    constexpr int HD = DPL * 32;
    extern __shared__ float smem_dyn[];
    __shared__ int s_last;
    pdl_launch_dependents();
    pdl_wait();
    const int pos = *pos_dev, seq_len = pos + 1;
    int split_len = (seq_len + n_splits - 1) / n_splits;
    if (split_len < min_split) split_len = min_split;      // short contexts: few, reasonably long slices
    const int used = (seq_len + split_len - 1) / split_len;
    const int head0 = blockIdx.x * GC, split = blockIdx.y;
    const int per_kv = n_heads / n_kv, kv_head = head0 / per_kv;
    const int k_begin = split * split_len, k_end = min(seq_len, k_begin + split_len);
    if (k_begin >= k_end) return;
    NewToken nt;
    nt.k = k_new; nt.v = v_new; nt.pos = pos; nt.theta = theta; nt.freq_scale = freq_scale;
    if (head0 % per_kv == 0) { nt.kc_w = kc; nt.vc_w = vc; }      // one CTA per KV head stores the new row (the split that holds pos)
    if (used == 1 && !xq_out && tickets) {                 // short context: one slice per head group, normalised output, no merge
        attend_group<DPL, GC>(out + (size_t)head0 * HD, q + (size_t)head0 * HD, kc, vc, kv_head, n_kv, k_begin, k_end, scale, nullptr, nullptr,
                              smem_dyn, nt);
        return;
    }
    float* stage = smem_dyn + (size_t)GC * (k_end - k_begin) + (size_t)AW * GC * HD;   // [GC][HD] + [GC][2]
    attend_group<DPL, GC>(nullptr, q + (size_t)head0 * HD, kc, vc, kv_head, n_kv, k_begin, k_end, scale, stage, stage + GC * HD,
                          smem_dyn, nt);
    __syncthreads();
    float* ml_all = scratch + (size_t)n_heads * n_splits * HD;
    for (int idx = threadIdx.x; idx < GC * HD; idx += AW * 32) {
        const int g = idx / HD, d = idx - g * HD;
        scratch[((size_t)(head0 + g) * n_splits + split) * HD + d] = stage[idx];
    }
    if (threadIdx.x < 2 * GC) {
        const int g = threadIdx.x >> 1;
        ml_all[((size_t)(head0 + g) * n_splits + split) * 2 + (threadIdx.x & 1)] = stage[GC * HD + threadIdx.x];
    }
    if (!tickets) return;                                  // two-launch form: decode_combine_kernel merges (attention_decode_rope_dyn)
    // ---- last arriver of the head group merges the splits ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(tickets + blockIdx.x, 1u);
        s_last = ((int)prev + 1 == used);
        if (s_last) tickets[blockIdx.x] = 0;                   // ready for the next launch (all arrivals are in)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < GC; g += AW) {
        const int h = head0 + g;
        const float* ml = ml_all + (size_t)h * n_splits * 2;
        float m = -FLT_MAX;
        for (int i = 0; i < used; i++) m = fmaxf(m, __ldcg(ml + 2 * i));
        float l = 0.f;
        for (int i = 0; i < used; i++) l += __ldcg(ml + 2 * i + 1) * expf(__ldcg(ml + 2 * i) - m);
        const float inv = (l > 0.f) ? 1.0f / l : 0.f;
        for (int d = lane; d < HD; d += 32) {
            float o = 0.f;
            for (int i = 0; i < used; i++) o += __ldcg(scratch + ((size_t)h * n_splits + i) * HD + d) * expf(__ldcg(ml + 2 * i) - m);
            const float v = o * inv;
            out[(size_t)h * HD + d] = v;
            if (xq_out) {
                const int K = n_heads * HD, e = h * HD + d;
                int q1, q2, q3;
                float sc, s16;
                quantize_lane32(v, q1, q2, q3, sc, s16);
                const int se = (int)xq_swizzle((uint32_t)e);
                xq_out[se] = (int8_t)q1;
                xq_out[K + se] = (int8_t)q2;
                xq_out[2 * K + se] = (int8_t)q3;
                float* xscale = reinterpret_cast<float*>(xq_out + 3 * (size_t)K);
                if (lane == 0) xscale[e >> 5] = sc;
                if ((lane & 15) == 0) xscale[K / 32 + (e >> 4)] = s16;
            }
        }
    }
}

template <int DPL, int GC>
__global__ void __launch_bounds__(AW * 32) prefill_kernel(float* __restrict__ out, const float* __restrict__ Q,
                                                          const __half* __restrict__ kc, const __half* __restrict__ vc,
                                                          int start_pos, int n_heads, int n_kv, float scale) {
    constexpr int HD = DPL * 32;
    extern __shared__ float smem_dyn[];
    const int head0 = blockIdx.x * GC, qi = blockIdx.y;
    const int kv_head = head0 / (n_heads / n_kv);
    const size_t qoff = ((size_t)qi * n_heads + head0) * HD;
    attend_group<DPL, GC>(out + qoff, Q + qoff, kc, vc, kv_head, n_kv, 0, start_pos + qi + 1, scale, nullptr, nullptr, smem_dyn);
}

float* g_scratch = nullptr;
size_t g_scratch_floats = 0;
std::mutex g_scratch_mu;
float* attn_scratch(size_t floats) {
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    if (floats > g_scratch_floats) {
        if (g_scratch) { cudaDeviceSynchronize(); cudaFree(g_scratch); }
        size_t want = floats < (1u << 20) ? (1u << 20) : floats;
        NT_CUDA_CHECK(cudaMalloc(&g_scratch, want * sizeof(float)));
        g_scratch_floats = want;
    }
    return g_scratch;
}

int pick_gc(int group) { return group % 8 == 0 ? 8 : group % 4 == 0 ? 4 : group % 2 == 0 ? 2 : 1; }

template <int DPL, int GC>
void launch_decode(float* out, const float* q, const __half* kc, const __half* vc, int seq_len, int n_heads, int n_kv,
                   float scale, cudaStream_t s) {
    constexpr int HD = DPL * 32;
    // split so that head-groups x splits covers the chip, with at least 32 keys per split
    int groups = n_heads / GC;
    int n_splits = 1;
    if (seq_len > 64) {
        n_splits = (2 * 148 + groups - 1) / groups;
        int max_by_len = (seq_len + 31) / 32;
        if (n_splits > max_by_len) n_splits = max_by_len;
        if (n_splits > MAX_SPLITS) n_splits = MAX_SPLITS;
        if (n_splits < 1) n_splits = 1;
    }
    int split_len = (seq_len + n_splits - 1) / n_splits;
    n_splits = (seq_len + split_len - 1) / split_len;
    size_t smem = ((size_t)GC * split_len + (size_t)AW * GC * HD + (size_t)GC * HD + 2 * GC) * sizeof(float);
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(decode_kernel<DPL, GC>, (int)(ATTN_MAX_DYN_SMEM), configured);
    NT_CHECK(smem <= (size_t)ATTN_MAX_DYN_SMEM, "attention_decode: context slice does not fit shared memory");
    float* scratch = nullptr;
    if (n_splits > 1) scratch = attn_scratch((size_t)n_heads * n_splits * (HD + 2));
    decode_kernel<DPL, GC><<<dim3(groups, n_splits), AW * 32, smem, s>>>(out, q, kc, vc, seq_len, n_heads, n_kv, scale,
                                                                        n_splits, split_len, scratch, nullptr);
    count_launch();
    if (n_splits > 1) {
        decode_combine_kernel<<<n_heads, 128, 0, s>>>(out, scratch, n_heads, HD, n_splits, seq_len, split_len, nullptr, nullptr);
        count_launch();
    }
}

// Graph-replayable decode: the context length (pos + 1) is read from device memory, the launch shape is fixed
// by max_seq, partials always go through caller-owned scratch (n_heads * n_splits * (HD + 2) floats).
template <int DPL, int GC>
void launch_decode_dyn(float* out, const float* q, const __half* kc, const __half* vc, const int* pos_dev, int max_seq,
                       int n_heads, int n_kv, float scale, float* scratch, int n_splits, int8_t* xq_out, cudaStream_t s) {
    constexpr int HD = DPL * 32;
    const int groups = n_heads / GC;
    const int max_split_len = std::max((max_seq + n_splits - 1) / n_splits, DYN_MIN_SPLIT);
    size_t smem = ((size_t)GC * max_split_len + (size_t)AW * GC * HD + (size_t)GC * HD + 2 * GC) * sizeof(float);
    NT_CHECK(smem <= (size_t)ATTN_MAX_DYN_SMEM, "attention_decode_dyn: context slice does not fit shared memory");
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(decode_kernel<DPL, GC>, (int)(ATTN_MAX_DYN_SMEM), configured);
    launch_k(decode_kernel<DPL, GC>, dim3(groups, n_splits), dim3(AW * 32), smem, s, out, q, kc, vc, 0, n_heads, n_kv, scale,
             n_splits, 0, scratch, pos_dev);
    launch_k(decode_combine_kernel, dim3(n_heads), dim3(128), 0, s, out, (const float*)scratch, n_heads, HD, n_splits, 0, 0, pos_dev, xq_out);
    count_launch(2);
}

template <int DPL, int GC>
void launch_decode_fused(float* out, const float* q, const float* k_new, const float* v_new, __half* kc, __half* vc, const int* pos_dev,
                         int max_seq, int n_heads, int n_kv, float theta, float freq_scale, float scale, float* scratch, int n_splits,
                         unsigned* tickets, int8_t* xq_out, cudaStream_t s) {
    constexpr int HD = DPL * 32;
    const int groups = n_heads / GC;
    int max_split_len = (max_seq + n_splits - 1) / n_splits;
    if (max_split_len < FUSED_MIN_SPLIT) max_split_len = FUSED_MIN_SPLIT;
    size_t smem = ((size_t)GC * max_split_len + (size_t)AW * GC * HD + (size_t)GC * HD + 2 * GC) * sizeof(float);
    NT_CHECK(smem <= (size_t)ATTN_MAX_DYN_SMEM, "attention_decode_fused: context slice does not fit shared memory");
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(decode_fused_kernel<DPL, GC>, (int)(ATTN_MAX_DYN_SMEM), configured);
    launch_k(decode_fused_kernel<DPL, GC>, dim3(groups, n_splits), dim3(AW * 32), smem, s, out, q, k_new, v_new, kc, vc, pos_dev, n_heads,
             n_kv, scale, theta, freq_scale, n_splits, FUSED_MIN_SPLIT, scratch, tickets, xq_out);
    count_launch();
}

// RoPE + KV write folded into the split-context decode kernel, merge as its own small launch (same slices as launch_decode_dyn):
// two launches instead of rope_kv_decode + decode + combine, without the one-launch form's serial chain or tickets.
template <int DPL, int GC>
void launch_decode_rope_dyn(float* out, const float* q, const float* k_new, const float* v_new, __half* kc, __half* vc, const int* pos_dev,
                            int max_seq, int n_heads, int n_kv, float theta, float freq_scale, float scale, float* scratch, int n_splits,
                            int8_t* xq_out, cudaStream_t s) {
    constexpr int HD = DPL * 32;
    const int groups = n_heads / GC;
    const int max_split_len = std::max((max_seq + n_splits - 1) / n_splits, DYN_MIN_SPLIT);
    size_t smem = ((size_t)GC * max_split_len + (size_t)AW * GC * HD + (size_t)GC * HD + 2 * GC) * sizeof(float);
    NT_CHECK(smem <= (size_t)ATTN_MAX_DYN_SMEM, "attention_decode_rope_dyn: context slice does not fit shared memory");
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(decode_fused_kernel<DPL, GC>, (int)(ATTN_MAX_DYN_SMEM), configured);
    launch_k(decode_fused_kernel<DPL, GC>, dim3(groups, n_splits), dim3(AW * 32), smem, s, out, q, k_new, v_new, kc, vc, pos_dev, n_heads,
             n_kv, scale, theta, freq_scale, n_splits, DYN_MIN_SPLIT, scratch, (unsigned*)nullptr, (int8_t*)nullptr);
    launch_k(decode_combine_kernel, dim3(n_heads), dim3(128), 0, s, out, (const float*)scratch, n_heads, HD, n_splits, 0, 0, pos_dev, xq_out);
    count_launch(2);
}

template <int DPL, int GC>
void launch_prefill(float* out, const float* Q, const __half* kc, const __half* vc, int seq_len, int start_pos, int n_heads,
                    int n_kv, float scale, cudaStream_t s) {
    constexpr int HD = DPL * 32;
    size_t smem = ((size_t)GC * (start_pos + seq_len) + (size_t)AW * GC * HD) * sizeof(float);
    NT_CHECK(smem <= (size_t)ATTN_MAX_DYN_SMEM, "attention_prefill: context does not fit shared memory");
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(prefill_kernel<DPL, GC>, (int)(ATTN_MAX_DYN_SMEM), configured);
    prefill_kernel<DPL, GC><<<dim3(n_heads / GC, seq_len), AW * 32, smem, s>>>(out, Q, kc, vc, start_pos, n_heads, n_kv, scale);
    count_launch();
}

#define NT_DISPATCH_ATTN(FN, ...)                                                         \
    do {                                                                                  \
        const int gc_ = pick_gc(n_heads / n_kv);                                          \
        const int dpl_ = hd / 32;                                                         \
        NT_CHECK(hd % 32 == 0 && (dpl_ == 2 || dpl_ == 4 || dpl_ == 8),                  \
                 "attention: head_dim must be 64, 128 or 256");                           \
        NT_CHECK(n_kv > 0 && n_heads % n_kv == 0, "attention: n_heads % n_kv_heads != 0"); \
        if (dpl_ == 4) {                                                                  \
            if (gc_ == 8) FN<4, 8>(__VA_ARGS__); else if (gc_ == 4) FN<4, 4>(__VA_ARGS__); \
            else if (gc_ == 2) FN<4, 2>(__VA_ARGS__); else FN<4, 1>(__VA_ARGS__);         \
        } else if (dpl_ == 2) {                                                           \
            if (gc_ == 8) FN<2, 8>(__VA_ARGS__); else if (gc_ == 4) FN<2, 4>(__VA_ARGS__); \
            else if (gc_ == 2) FN<2, 2>(__VA_ARGS__); else FN<2, 1>(__VA_ARGS__);         \
        } else {                                                                          \
            if (gc_ == 8) FN<8, 8>(__VA_ARGS__); else if (gc_ == 4) FN<8, 4>(__VA_ARGS__); \
            else if (gc_ == 2) FN<8, 2>(__VA_ARGS__); else FN<8, 1>(__VA_ARGS__);         \
        }                                                                                 \
    } while (0)

}  // namespace

void attention_decode(float* out, const float* q, const void* kc, const void* vc, int seq_len, int n_heads, int n_kv,
                      int hd, int max_seq, float scale, cudaStream_t s) {
    if (seq_len <= 0 || n_heads <= 0) return;
    const __half* k = static_cast<const __half*>(kc);
    const __half* v = static_cast<const __half*>(vc);
    NT_DISPATCH_ATTN(launch_decode, out, q, k, v, seq_len, n_heads, n_kv, scale, s);
}

int attention_decode_dyn_splits(int max_seq, int n_heads, int n_kv) {
    const int groups = n_heads / pick_gc(n_heads / n_kv);
    int n = (2 * 148 + groups - 1) / groups;
    int by_len = (max_seq + 31) / 32;
    if (n > by_len) n = by_len;
    if (n > MAX_SPLITS) n = MAX_SPLITS;
    return n < 1 ? 1 : n;
}
size_t attention_decode_dyn_scratch_floats(int max_seq, int n_heads, int n_kv, int hd) {
    return (size_t)n_heads * attention_decode_dyn_splits(max_seq, n_heads, n_kv) * (hd + 2);
}
void attention_decode_dyn(float* out, const float* q, const void* kc, const void* vc, const int* pos_dev, int max_seq,
                          int n_heads, int n_kv, int hd, float scale, float* scratch, void* xq_out, cudaStream_t s) {
    const __half* k = static_cast<const __half*>(kc);
    const __half* v = static_cast<const __half*>(vc);
    const int n_splits = attention_decode_dyn_splits(max_seq, n_heads, n_kv);
    if (xq_out) NT_CHECK((n_heads * hd) % 128 == 0, "attention_decode_dyn: n_heads * head_dim must be a multiple of 128 for the fused quantiser");
    NT_DISPATCH_ATTN(launch_decode_dyn, out, q, k, v, pos_dev, max_seq, n_heads, n_kv, scale, scratch, n_splits, static_cast<int8_t*>(xq_out), s);
}

int attention_decode_fused_tickets(int n_heads, int n_kv) { return n_heads / pick_gc(n_heads / n_kv); }
void attention_decode_fused(float* out, const float* q, const float* k, const float* v, void* kc, void* vc, const int* pos_dev,
                            int max_seq, int n_heads, int n_kv, int hd, float theta, float freq_scale, float scale, float* scratch,
                            unsigned* tickets, void* xq_out, cudaStream_t s) {
    __half* kh = static_cast<__half*>(kc);
    __half* vh = static_cast<__half*>(vc);
    const int n_splits = attention_decode_dyn_splits(max_seq, n_heads, n_kv);
    if (xq_out) NT_CHECK((n_heads * hd) % 128 == 0, "attention_decode_fused: n_heads * head_dim must be a multiple of 128 for the fused quantiser");
    NT_DISPATCH_ATTN(launch_decode_fused, out, q, k, v, kh, vh, pos_dev, max_seq, n_heads, n_kv, theta, freq_scale, scale, scratch, n_splits,
                     tickets, static_cast<int8_t*>(xq_out), s);
}

void attention_decode_rope_dyn(float* out, const float* q, const float* k, const float* v, void* kc, void* vc, const int* pos_dev,
                               int max_seq, int n_heads, int n_kv, int hd, float theta, float freq_scale, float scale, float* scratch,
                               void* xq_out, cudaStream_t s) {
    __half* kh = static_cast<__half*>(kc);
    __half* vh = static_cast<__half*>(vc);
    const int n_splits = attention_decode_dyn_splits(max_seq, n_heads, n_kv);
    if (xq_out) NT_CHECK((n_heads * hd) % 128 == 0, "attention_decode_rope_dyn: n_heads * head_dim must be a multiple of 128 for the fused quantiser");
    NT_DISPATCH_ATTN(launch_decode_rope_dyn, out, q, k, v, kh, vh, pos_dev, max_seq, n_heads, n_kv, theta, freq_scale, scale, scratch, n_splits,
                     static_cast<int8_t*>(xq_out), s);
}

void attention_prefill(float* out, const float* Q, const void* kc, const void* vc, int seq_len, int start_pos, int n_heads,
                       int n_kv, int hd, int max_seq, float scale, cudaStream_t s) {
    if (seq_len <= 0 || n_heads <= 0) return;
    if (attention_prefill_mma_supported(seq_len, n_heads, n_kv, hd)) {       // prompt-sized chunks: tiled tensor-core kernel
        attention_prefill_mma(out, Q, kc, vc, seq_len, start_pos, n_heads, n_kv, hd, max_seq, scale, s);
        return;
    }
    const __half* k = static_cast<const __half*>(kc);
    const __half* v = static_cast<const __half*>(vc);
    NT_DISPATCH_ATTN(launch_prefill, out, Q, k, v, seq_len, start_pos, n_heads, n_kv, scale, s);
}

}}  // namespace nt::b200
