// sample.cu — repeat penalty + temperature + top-k + top-p + inverse-CDF draw on the GPU.
//
// Replaces, for 0 < top_k <= 1024 < vocab, the reference's CPU sampler (Sampler::apply_repeat_penalty + Sampler::sample,
// src/inference/sampler.cpp:30-117), which copies 513 KB of logits to the host and runs a partial_sort over 128 k candidates
// per token (0.1-1 ms: comparable to a whole 8B token on a B200).  Here the token id is the only thing that
// leaves the GPU.  The arithmetic follows the reference step by step so that the draw is reproducible against it:
//   * penalty: sequential over the window (a token that occurs twice is penalised twice), IEEE division;
//   * candidates: logit / temperature (IEEE), the k largest (radix select on order-preserving keys, ties at the threshold
//     resolved towards the lower token id — the reference's std::partial_sort leaves that order unspecified), sorted by
//     (value desc, id asc) with a bitonic network in shared memory;
//   * softmax / top-p / renormalisation / cumulative draw: one thread, the same sequence of float additions and IEEE
//     divisions as the reference loops; exp() is evaluated in double and rounded to float, which reproduces glibc's
//     (correctly rounded in practice) expf where the fast-math ex2.approx path would not;
//   * the uniform variate r comes from the host's std::mt19937 exactly as in the reference (Sampler::draw in engine/text.cpp).
// STATUS: written after round 1's GPU budget was spent; verified on the CPU emulator (tests/cusim, tests/test_sample_sim.py)
// against the reference's own sampler; opt-in (GenerateConfig::gpu_sampler / NT_B200_GPU_SAMPLER=1) until it has run on hardware.
#include "kernels_internal.h"
#include <cstdint>

namespace nt { namespace b200 {

namespace {

constexpr int ST = 1024;                 // threads of the single CTA
constexpr int MAX_K = 1024;

struct SampleParams {
    float* logits;                       // [n] device; the repeat penalty is applied in place
    int n;
    float temperature, top_p, penalty, r;
    int top_k;
    const int* recent;                   // device: the window's token ids, oldest first
    int n_recent;
    int* out;                            // device: [0] sampled token id
};

__device__ __forceinline__ uint32_t order_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct SampleShared {
    unsigned hist[256];
    unsigned long long cand[MAX_K];      // (key << 32) | (0xFFFFFFFF - token id): descending sort = value desc, id asc
    float prob[MAX_K];
    unsigned scan[ST];
    unsigned prefix, mask, k_rem, eq_count, count, running;
};

__global__ void __launch_bounds__(ST, 1) sample_topk_kernel(const SampleParams p) {
    __shared__ SampleShared S;
    const int tid = threadIdx.x;
    const int n = p.n, k = p.top_k;

    // ---- repeat penalty (sampler.cpp:30-45): sequential, in place ----
    if (tid == 0 && p.penalty > 1.0f) {
        for (int i = 0; i < p.n_recent; i++) {
            const int t = p.recent[i];
            if (t < 0 || t >= n) continue;
            const float l = p.logits[t];
            p.logits[t] = (l > 0.f) ? __fdiv_rn(l, p.penalty) : __fmul_rn(l, p.penalty);
        }
    }
    if (tid == 0) { S.prefix = 0; S.mask = 0; S.k_rem = (unsigned)k; S.count = 0; S.running = 0; }
    __syncthreads();

    // ---- radix select: key of the k-th largest candidate, most significant byte first ----
    for (int pass = 3; pass >= 0; pass--) {
        if (tid < 256) S.hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = S.prefix, mask = S.mask;
        for (int i = tid; i < n; i += ST) {
            const uint32_t key = order_key(__fdiv_rn(p.logits[i], p.temperature));
            if ((key & mask) == prefix) atomicAdd(&S.hist[(key >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned cum = 0, rem = S.k_rem;
            int d = 255;
            for (; d > 0; d--) {
                if (cum + S.hist[d] >= rem) break;
                cum += S.hist[d];
            }
            S.k_rem = rem - cum;                             // how many candidates with this digit (finally: this key) are needed
            S.eq_count = S.hist[d];
            S.prefix = prefix | ((unsigned)d << (8 * pass));
            S.mask = mask | (255u << (8 * pass));
        }
        __syncthreads();
    }
    const uint32_t tkey = S.prefix;                          // key of the k-th largest value
    const unsigned need_eq = S.k_rem, n_gt = (unsigned)k - need_eq;
    const bool take_all_eq = (S.eq_count == need_eq);

    // ---- collect: everything above the threshold, plus need_eq candidates equal to it (lowest ids first) ----
    for (int i = tid; i < n; i += ST) {
        const uint32_t key = order_key(__fdiv_rn(p.logits[i], p.temperature));
        if (key > tkey || (take_all_eq && key == tkey)) {
            const unsigned pos = atomicAdd(&S.count, 1u);
            if (pos < (unsigned)MAX_K) S.cand[pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        }
    }
    __syncthreads();
    if (!take_all_eq) {
        // ties straddle the cut: walk the ids in order, a block-wide scan per chunk ranks the equal ones
        for (int base = 0; base < n; base += ST) {
            const int i = base + tid;
            const unsigned flag = (i < n && order_key(__fdiv_rn(p.logits[i], p.temperature)) == tkey) ? 1u : 0u;
            S.scan[tid] = flag;
            __syncthreads();
            for (int off = 1; off < ST; off <<= 1) {         // Hillis-Steele inclusive scan
                const unsigned add = (tid >= off) ? S.scan[tid - off] : 0u;
                __syncthreads();
                S.scan[tid] += add;
                __syncthreads();
            }
            const unsigned rank = S.running + S.scan[tid] - flag;
            if (flag && rank < need_eq)
                S.cand[n_gt + rank] = ((unsigned long long)tkey << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
            __syncthreads();
            if (tid == 0) S.running += S.scan[ST - 1];
            __syncthreads();
            if (S.running >= need_eq) break;                 // CTA-uniform
        }
    }
    // ---- bitonic sort, descending ----
    int P = 1;
    while (P < k) P <<= 1;
    if (tid >= k && tid < P) S.cand[tid] = 0ull;
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int j = tid ^ stride;
            if (tid < P && j > tid) {
                const bool desc = (tid & size) == 0;
                const unsigned long long a = S.cand[tid], b = S.cand[j];
                if ((a < b) == desc) { S.cand[tid] = b; S.cand[j] = a; }
            }
            __syncthreads();
        }
    }
    // ---- exp(logit - max) per candidate (sampler.cpp:72-77); double exp rounded to float ----
    if (tid < k) {
        const float mx = key_value((uint32_t)(S.cand[0] >> 32));
        const float v = key_value((uint32_t)(S.cand[tid] >> 32));
        S.prob[tid] = (float)exp((double)__fsub_rn(v, mx));
    }
    __syncthreads();
    // ---- normalise, top-p, renormalise, draw: the reference's loops, one thread ----
    if (tid == 0) {
        float sum = 0.f;
        for (int i = 0; i < k; i++) sum = __fadd_rn(sum, S.prob[i]);
        for (int i = 0; i < k; i++) S.prob[i] = __fdiv_rn(S.prob[i], sum);
        int m = k;
        if (p.top_p < 1.0f && p.top_p > 0.0f) {
            float cum = 0.f;
            for (int i = 0; i < k; i++) {
                cum = __fadd_rn(cum, S.prob[i]);
                if (cum >= p.top_p) { m = i + 1; break; }
            }
            float s2 = 0.f;
            for (int i = 0; i < m; i++) s2 = __fadd_rn(s2, S.prob[i]);
            for (int i = 0; i < m; i++) S.prob[i] = __fdiv_rn(S.prob[i], s2);
        }
        int pick = m - 1;
        float cum = 0.f;
        for (int i = 0; i < m; i++) {
            cum = __fadd_rn(cum, S.prob[i]);
            if (p.r <= cum) { pick = i; break; }
        }
        p.out[0] = (int)(0xFFFFFFFFu - (unsigned)(S.cand[pick] & 0xFFFFFFFFull));
    }
}

}  // namespace

bool sample_topk_supported(int n, float temperature, int top_k) { return temperature > 0.0f && top_k > 0 && top_k <= MAX_K && top_k < n; }

#ifndef NT_CUSIM
bool sample_topk(float* logits_dev, int n, float temperature, int top_k, float top_p, float repeat_penalty, const int* recent_dev,
                 int n_recent, float r, int* out_dev, cudaStream_t s) {
    if (!sample_topk_supported(n, temperature, top_k)) return false;
    SampleParams p{logits_dev, n, temperature, top_p, repeat_penalty, r, top_k, recent_dev, n_recent, out_dev};
    sample_topk_kernel<<<1, ST, 0, s>>>(p);
    count_launch();
    return true;
}
#endif

}}  // namespace nt::b200
