// gemv_kquant.cu — dequant-fused GEMV over GGUF Q4_K / Q5_K / Q6_K super-blocks for sm_100a.
//
// Replaces reference kernels K3/K4/K5 (src/cuda/gemm.cu:158-470) behind launch_gemv (gemm.cu:748-805).
// The reference walks one warp per row with per-lane byte loads, an F32 convert + 2 FMAs per weight and
// x re-read from shared memory for every row.  On B200 (7.3 TB/s reads, ~48 Q4_K weights/clk/SM) that
// instruction stream is ~2x over the issue budget, so this kernel is organised differently:
//
//   * weights: each warp owns a private ring of TMA bulk copies (cp.async.bulk + mbarrier): one stage =
//     4 rows x 16 super-blocks (>= 2304 B per copy), filled by the warp's lane 0 and consumed by the
//     same warp, so there is no CTA-wide barrier in the streaming loop and HBM requests stay in flight
//     across row-group boundaries;
//   * activations: block-scaled 3-term int8 (kernels_internal.h "xq"), staged once per CTA in shared
//     memory (XOR-swizzled so the lane -> half-block mapping is bank-conflict free) and re-used for the
//     4 rows of a stage from registers;
//   * math: IDP.4A on the 4/5/6-bit codes (exact integer partial sums), one F32 scale per 16/32 weights,
//     6-bit scale unpack amortised over a 128-weight half super-block per lane;
//   * reduction: a 4-row transpose-reduce (6 shuffles) once per row-group.
// The kernel is specialised per block format (compile-time block size / row pitch: no address IMADs) and
// the CTA width (4..8 warps) is chosen per launch to minimise rounds x warps (tail quantisation).
// No tensor cores: the path is HBM-bound (BASELINE.json north_star).
#include "kernels_internal.h"
#include "ring.cuh"
#include "xquant.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace nt { namespace b200 {

namespace {

constexpr int RG = 4;            // rows per warp stage
constexpr int BS = 16;           // super-blocks per stage chunk (two lanes per super-block)
constexpr int MAX_MATS = 3;

template <int FMT> struct Fmt;
template <> struct Fmt<0> { static constexpr int BLK = 144; };   // Q4_K
template <> struct Fmt<1> { static constexpr int BLK = 176; };   // Q5_K
template <> struct Fmt<2> { static constexpr int BLK = 210; };   // Q6_K

struct KqMat {
    const uint8_t* W;
    float* y;
    int out;
    int groups;        // ceil(out / RG)
    int fmt;           // 0 Q4_K, 1 Q5_K, 2 Q6_K
    long long row_pitch;
};
struct KqParams {
    KqMat mat[MAX_MATS];
    int n_mat;
    int K, NB, NC;             // elements, super-blocks per row, chunks per row
    const int8_t* xq;          // pre-quantised activations, or null:
    const float* x_f32;        //   F32 activations quantised in the prologue,
    const float* norm_w;       //   optionally RMS-normalised first (x * rsqrt(mean(x^2) + eps) * norm_w)
    float eps;
    int total_groups;          // SWIGLU: groups of mat[0]
    int n_seg;                 // segments (matrices) per task: 2 for SWIGLU else 1
    int epilogue;
    int stages;                // ring depth per warp
};

__device__ __forceinline__ float h2f(uint32_t h16) { return __half2float(__ushort_as_half((unsigned short)h16)); }
__device__ __forceinline__ int combine3(int s0, int s1, int s2) { return (s0 * 128 + s1) * 128 + s2; }

// Stage cursor: (task, segment, chunk) advanced without divisions.
struct Cursor {
    int g;        // global row-group index of the current task (gw + task * nw)
    int seg, chunk;
    int mi, gl;   // matrix index and row-group inside that matrix
};

// One stage (RG rows x up to BS super-blocks) of format FMT: this lane's half super-block against its x terms.
template <int FMT>
__device__ __forceinline__ void process_stage(const uint8_t* __restrict__ slot_base, int blk, int h, uint32_t hb,
                                              const uint8_t* __restrict__ xs, int K, const float* __restrict__ xscale,
                                              const float* __restrict__ xsum16, float (&acc)[RG]) {
    constexpr int BLK = Fmt<FMT>::BLK;
    constexpr int ROWP = BS * BLK;            // row pitch inside a stage slot
    const uint8_t* base = slot_base + blk * BLK;
    const uint8_t* xh = xs + hb * 128u;
    const uint32_t sw = hb & 7u;
    if (FMT <= 1) {
        // ---------------- Q4_K / Q5_K ----------------
        constexpr int QS = (FMT == 0) ? 16 : 48;
        uint32_t sc4[RG], m4[RG];
        float d[RG], dmin[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) {
            const int4 hd = *reinterpret_cast<const int4*>(base + r * ROWP);
            const uint32_t w0 = hd.y, w1 = hd.z, w2 = hd.w;
            d[r] = h2f((uint32_t)hd.x & 0xFFFFu);
            dmin[r] = h2f((uint32_t)hd.x >> 16);
            const uint32_t sa = w0 & 0x3F3F3F3Fu, ma = w1 & 0x3F3F3F3Fu;
            const uint32_t sb = (w2 & 0x0F0F0F0Fu) | ((w0 >> 2) & 0x30303030u);
            const uint32_t mb = ((w2 >> 4) & 0x0F0F0F0Fu) | ((w1 >> 2) & 0x30303030u);
            sc4[r] = h ? sb : sa;
            m4[r] = h ? mb : ma;
        }
        float A[RG] = {0.f, 0.f, 0.f, 0.f}, B[RG] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c2 = 0; c2 < 2; c2++) {
            int xr[3][16];
#pragma unroll
            for (int pl = 0; pl < 3; pl++) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int4 v = *reinterpret_cast<const int4*>(xh + pl * K + (((uint32_t)(c2 * 4 + i) ^ sw) << 4));
                    xr[pl][4 * i + 0] = v.x; xr[pl][4 * i + 1] = v.y; xr[pl][4 * i + 2] = v.z; xr[pl][4 * i + 3] = v.w;
                }
            }
            const int b32 = hb * 4 + c2 * 2, b16 = hb * 8 + c2 * 4;
            const float sx_lo = xscale[b32], sx_hi = xscale[b32 + 1];
            const float sum_lo = xsum16[b16] + xsum16[b16 + 1], sum_hi = xsum16[b16 + 2] + xsum16[b16 + 3];
            const int cg = 2 * h + c2;
            // fetch the codes of all RG rows first (independent 16-byte loads in flight together)
            int4 qa[RG], qb[RG], ha[RG], hbv[RG];
#pragma unroll
            for (int r = 0; r < RG; r++) {
                const uint8_t* qp = base + r * ROWP + QS + h * 64 + c2 * 32;
                qa[r] = *reinterpret_cast<const int4*>(qp);
                qb[r] = *reinterpret_cast<const int4*>(qp + 16);
                if (FMT == 1) {
                    ha[r] = *reinterpret_cast<const int4*>(base + r * ROWP + 16);
                    hbv[r] = *reinterpret_cast<const int4*>(base + r * ROWP + 32);
                }
            }
#pragma unroll
            for (int r = 0; r < RG; r++) {
                const uint32_t q[8] = {(uint32_t)qa[r].x, (uint32_t)qa[r].y, (uint32_t)qa[r].z, (uint32_t)qa[r].w,
                                       (uint32_t)qb[r].x, (uint32_t)qb[r].y, (uint32_t)qb[r].z, (uint32_t)qb[r].w};
                uint32_t qh[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (FMT == 1) {
                    qh[0] = ha[r].x; qh[1] = ha[r].y; qh[2] = ha[r].z; qh[3] = ha[r].w;
                    qh[4] = hbv[r].x; qh[5] = hbv[r].y; qh[6] = hbv[r].z; qh[7] = hbv[r].w;
                }
                int l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    uint32_t lo = q[w] & 0x0F0F0F0Fu;
                    uint32_t hi = (q[w] >> 4) & 0x0F0F0F0Fu;
                    if (FMT == 1) {
                        const uint32_t t = qh[w] >> (2 * cg);     // bit0 -> low sub-block, bit1 -> high sub-block
                        lo |= (t << 4) & 0x10101010u;
                        hi |= (t << 3) & 0x10101010u;
                    }
                    l0 = dp4a_us(lo, xr[0][w], l0); l1 = dp4a_us(lo, xr[1][w], l1); l2 = dp4a_us(lo, xr[2][w], l2);
                    h0 = dp4a_us(hi, xr[0][8 + w], h0); h1 = dp4a_us(hi, xr[1][8 + w], h1); h2 = dp4a_us(hi, xr[2][8 + w], h2);
                }
                const float flo = (float)combine3(l0, l1, l2) * sx_lo;
                const float fhi = (float)combine3(h0, h1, h2) * sx_hi;
                const uint32_t s2 = sc4[r] >> (16 * c2), m2 = m4[r] >> (16 * c2);
                A[r] = fmaf((float)(s2 & 0xFFu), flo, fmaf((float)((s2 >> 8) & 0xFFu), fhi, A[r]));
                B[r] = fmaf((float)(m2 & 0xFFu), sum_lo, fmaf((float)((m2 >> 8) & 0xFFu), sum_hi, B[r]));
            }
        }
#pragma unroll
        for (int r = 0; r < RG; r++) acc[r] += d[r] * A[r] - dmin[r] * B[r];
    } else {
        // ---------------- Q6_K (210-byte blocks: 2-byte aligned, realigned with PRMT) ----------------
        const uint32_t mis = (uint32_t)(blk & 1) * 2u;           // (blk * 210) & 2
        const uint32_t sel = mis ? 0x5432u : 0x3210u;
        const uint8_t* ab = base - mis;                            // 4-byte aligned view of the block
        uint32_t scw[RG][2];
        float d[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) {
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 192 + 8 * h);
            const uint32_t a0 = sp[0], a1 = sp[1], a2 = sp[2];
            scw[r][0] = __byte_perm(a0, a1, sel);
            scw[r][1] = __byte_perm(a1, a2, sel);
            const uint32_t dw = *reinterpret_cast<const uint32_t*>(ab + r * ROWP + 208);
            d[r] = h2f(mis ? (dw >> 16) : (dw & 0xFFFFu));
        }
        float A[RG] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
            int xr[3][16];
            float sx[4], c32[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int pl = 0; pl < 3; pl++) {
                    const int4 v = *reinterpret_cast<const int4*>(xh + pl * K + (((uint32_t)(2 * j + kk) ^ sw) << 4));
                    xr[pl][4 * j + 0] = v.x; xr[pl][4 * j + 1] = v.y; xr[pl][4 * j + 2] = v.z; xr[pl][4 * j + 3] = v.w;
                }
                sx[j] = xscale[hb * 4 + j];
                c32[j] = 32.0f * xsum16[hb * 8 + 2 * j + kk];
            }
#pragma unroll
            for (int r = 0; r < RG; r++) {
                const uint32_t* pa = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 64 * h + 16 * kk);          // ql[l]
                const uint32_t* pb = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 64 * h + 32 + 16 * kk);     // ql[l+32]
                const uint32_t* ph = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 128 + 32 * h + 16 * kk);    // qh[l]
                uint32_t ra[5], rb[5], rh[5];
#pragma unroll
                for (int i = 0; i < 5; i++) { ra[i] = pa[i]; rb[i] = pb[i]; rh[i] = ph[i]; }
                int s[4][3] = {};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t qa = __byte_perm(ra[i], ra[i + 1], sel), qb = __byte_perm(rb[i], rb[i + 1], sel);
                    const uint32_t hh = __byte_perm(rh[i], rh[i + 1], sel);
                    const uint32_t q1 = (qa & 0x0F0F0F0Fu) | ((hh << 4) & 0x30303030u);
                    const uint32_t q2 = (qb & 0x0F0F0F0Fu) | ((hh << 2) & 0x30303030u);
                    const uint32_t q3 = ((qa >> 4) & 0x0F0F0F0Fu) | (hh & 0x30303030u);
                    const uint32_t q4 = ((qb >> 4) & 0x0F0F0F0Fu) | ((hh >> 2) & 0x30303030u);
#pragma unroll
                    for (int pl = 0; pl < 3; pl++) {
                        s[0][pl] = dp4a_us(q1, xr[pl][0 + i], s[0][pl]);
                        s[1][pl] = dp4a_us(q2, xr[pl][4 + i], s[1][pl]);
                        s[2][pl] = dp4a_us(q3, xr[pl][8 + i], s[2][pl]);
                        s[3][pl] = dp4a_us(q4, xr[pl][12 + i], s[3][pl]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int b = 2 * j + kk;                      // scale index inside the half
                    const int sc = (int)(signed char)((scw[r][b >> 2] >> (8 * (b & 3))) & 0xFFu);
                    // sum over 16 weights of sc * (q - 32) * x = sc * (S * sx - 32 * sum16)
                    A[r] = fmaf((float)sc, fmaf((float)combine3(s[j][0], s[j][1], s[j][2]), sx[j], -c32[j]), A[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RG; r++) acc[r] = fmaf(d[r], A[r], acc[r]);
    }
}

__host__ __device__ constexpr int max_blk(int mask) { return (mask & 4) ? 210 : (mask & 2) ? 176 : 144; }

// MASK: bit f set <=> matrices of format f may appear in this launch (mixed Q4_K_M projections share one launch).
template <int MASK, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) gemv_kq_kernel(const __grid_constant__ KqParams p) {
    constexpr int SLOT = RG * BS * max_blk(MASK);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float red[32];
    const int K = p.K;
    // smem carve-up: [x planes 3K][scale K/32 f32][sum16 K/16 f32][pad to 128][rings][mbarriers]
    uint8_t* xs = smem;
    float* xscale = reinterpret_cast<float*>(smem + 3 * (size_t)K);
    float* xsum16 = xscale + K / 32;
    const size_t ring_off = (3 * (size_t)K + (size_t)(K / 32) * 4 + (size_t)(K / 16) * 4 + 127) & ~(size_t)127;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    uint8_t* ring = smem + ring_off + (size_t)warp * stages * SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_off + (size_t)WARPS * stages * SLOT) + warp * stages;

    const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
    const int n_tasks = (gw < p.total_groups) ? (p.total_groups - gw + nw - 1) / nw : 0;
    const int NC = p.NC, NB = p.NB, n_seg = p.n_seg;
    const int n_stages_total = n_tasks * n_seg * NC;

    auto locate = [&](Cursor& c) {            // matrix lookup for the cursor's task
        if (n_seg == 2) { c.mi = c.seg; c.gl = c.g; return; }
        int g = c.g, mi = 0;
        while (mi + 1 < p.n_mat && g >= p.mat[mi].groups) { g -= p.mat[mi].groups; mi++; }
        c.mi = mi; c.gl = g;
    };
    auto advance = [&](Cursor& c) {
        if (++c.chunk == NC) {
            c.chunk = 0;
            if (++c.seg == n_seg) { c.seg = 0; c.g += nw; }
            locate(c);
        }
    };
    auto issue = [&](const Cursor& c, int slot) {   // lane 0: TMA copies of one stage into ring slot
        const KqMat& m = p.mat[c.mi];
        const int nbc = min(BS, NB - c.chunk * BS);
        const int blkb = (MASK == 1) ? 144 : (MASK == 2) ? 176 : (MASK == 4) ? 210 : (m.fmt == 0 ? 144 : m.fmt == 1 ? 176 : 210);
        // copy size rounded up to 16 B (a 210-byte Q6_K tail may spill into row padding; checked on the host)
        const uint32_t bytes = ((uint32_t)(nbc * blkb) + 15u) & ~15u;
        uint64_t* bar = bars + slot;
        mbar_expect_tx(bar, bytes * RG);
        uint8_t* dst = ring + (size_t)slot * SLOT;
        const uint8_t* src = m.W + (long long)c.chunk * (BS * blkb);
#pragma unroll
        for (int r = 0; r < RG; r++) {
            const int row = min(c.gl * RG + r, m.out - 1);
            bulk_g2s(dst + r * (BS * blkb), src + (long long)row * m.row_pitch, bytes, bar);
        }
    };

    uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + ring_off + (size_t)WARPS * stages * SLOT) + WARPS * stages;
    if (lane == 0) {
        for (int s = 0; s < stages; s++) mbar_init(bars + s, 1);
        if (warp == 0) mbar_init(xbar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    Cursor pc{gw, 0, 0, 0, 0};                 // producer cursor (runs `stages` ahead of the consumer)
    locate(pc);
    int issued = 0;
    // Weights do not depend on the previous kernel: start streaming before touching x.
    if (lane == 0) {
        for (; issued < stages && issued < n_stages_total; issued++) { issue(pc, issued); advance(pc); }
    }
    pdl_wait();   // no-op unless launched with programmatic stream serialization

    if (p.x_f32) {
        // ---- fused prologue: (RMSNorm +) block-scaled int8x3 quantisation straight into shared memory ----
        // Replaces the reference's separate rmsnorm launch (rmsnorm.cu:17-70) for the consumer GEMV; every CTA
        // recomputes it from the L2-resident hidden state while its first weight stages are in flight.
        float rms_inv = 1.0f;
        if (p.norm_w) {
            float ss = 0.f;
            for (int i = threadIdx.x; i < K; i += WARPS * 32) { const float v = p.x_f32[i]; ss += v * v; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
            if (lane == 0) red[warp] = ss;
            __syncthreads();
            float t = (lane < WARPS) ? red[lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            const float mean_sq = t / K;
            rms_inv = rsqrtf(mean_sq + p.eps);
        }
        // batches of PB 32-element blocks per warp: the PB loads are issued together (one L2 latency per batch)
        constexpr int PB = 8;
        const int nblk = K / 32;
        for (int b0 = warp; b0 < nblk; b0 += WARPS * PB) {
            float v[PB], w[PB];
#pragma unroll
            for (int j = 0; j < PB; j++) {
                const int b = b0 + j * WARPS;
                const int e = min(b, nblk - 1) * 32 + lane;
                v[j] = p.x_f32[e];
                w[j] = p.norm_w ? p.norm_w[e] : 1.0f;
            }
#pragma unroll
            for (int j = 0; j < PB; j++) {
                const int b = b0 + j * WARPS;
                if (b >= nblk) break;                          // warp-uniform
                const int e = b * 32 + lane;
                const float x = p.norm_w ? v[j] * rms_inv * w[j] : v[j];
                int q1, q2, q3;
                float sc, s16;
                quantize_lane32(x, q1, q2, q3, sc, s16);
                const uint32_t se = xq_swizzle((uint32_t)e);
                xs[se] = (uint8_t)q1;
                xs[K + se] = (uint8_t)q2;
                xs[2 * K + se] = (uint8_t)q3;
                if (lane == 0) xscale[b] = sc;
                if ((lane & 15) == 0) xsum16[2 * b + (lane >> 4)] = s16;
            }
        }
        __syncthreads();
    } else {
        // ---- pre-quantised xq: the producer kernel already wrote the planes in the swizzled order, so staging is a
        //      plain TMA bulk copy (<= 32 KB pieces) tracked by one CTA-level mbarrier: no per-thread loads/stores ----
        if (threadIdx.x == 0) {
            const uint32_t total = (uint32_t)(3 * (size_t)K + (size_t)(K / 32) * 4 + (size_t)(K / 16) * 4);
            mbar_expect_tx(xbar, total);
            for (uint32_t off = 0; off < total; off += 32768u) {
                const uint32_t n = min(32768u, total - off);
                bulk_g2s(xs + off, p.xq + off, n, xbar);
            }
        }
        mbar_wait(xbar, 0);
    }
    pdl_launch_dependents();

    const int blk = lane >> 1, h = lane & 1;
    float acc[RG] = {0.f, 0.f, 0.f, 0.f};
    float gate_keep = 0.f;
    Cursor cc{gw, 0, 0, 0, 0};
    locate(cc);
    int slot = 0;
    uint32_t parity = 0;

    for (int s = 0; s < n_stages_total; s++) {
        const int nbc = min(BS, NB - cc.chunk * BS);
        mbar_wait(bars + slot, parity);

        if (blk < nbc) {
            const uint8_t* slot_base = ring + (size_t)slot * SLOT;
            const uint32_t hb = (uint32_t)((cc.chunk * BS + blk) * 2 + h);      // global half-block index
            const int fmt = (MASK == 1) ? 0 : (MASK == 2) ? 1 : (MASK == 4) ? 2 : p.mat[cc.mi].fmt;
            if ((MASK & 1) && fmt == 0) process_stage<0>(slot_base, blk, h, hb, xs, K, xscale, xsum16, acc);
            if ((MASK & 2) && fmt == 1) process_stage<1>(slot_base, blk, h, hb, xs, K, xscale, xsum16, acc);
            if ((MASK & 4) && fmt == 2) process_stage<2>(slot_base, blk, h, hb, xs, K, xscale, xsum16, acc);
        }
        __syncwarp();
        if (lane == 0 && issued < n_stages_total) { issue(pc, slot); advance(pc); }
        issued++;

        if (cc.chunk == NC - 1) {
            // ---- 4-row transpose-reduce: lanes with (lane & 7) == 0 end up holding one row each ----
            const bool b4 = lane & 16, b3 = lane & 8;
            const float s0 = b4 ? acc[0] : acc[2], s1 = b4 ? acc[1] : acc[3];
            float k0 = b4 ? acc[2] : acc[0], k1 = b4 ? acc[3] : acc[1];
            k0 += __shfl_xor_sync(0xFFFFFFFFu, s0, 16);
            k1 += __shfl_xor_sync(0xFFFFFFFFu, s1, 16);
            const float sv = b3 ? k0 : k1;
            float kv = b3 ? k1 : k0;
            kv += __shfl_xor_sync(0xFFFFFFFFu, sv, 8);
            kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 4);
            kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 2);
            kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 1);
            const KqMat& m = p.mat[cc.mi];
            const int row = cc.gl * RG + (b4 ? 2 : 0) + (b3 ? 1 : 0);
            if ((lane & 7) == 0) {
                if (p.epilogue == GEMV_SWIGLU) {
                    if (cc.seg == 0) {
                        gate_keep = kv;
                    } else if (row < m.out) {
                        // silu(g) * u with the reference's fast-math expression (gemm.cu:713-725)
                        const float gv = gate_keep;
                        p.mat[0].y[row] = __fdividef(gv, 1.0f + __expf(-gv)) * kv;
                    }
                } else if (row < m.out) {
                    if (p.epilogue == GEMV_ADD) m.y[row] += kv; else m.y[row] = kv;
                }
            }
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        }
        advance(cc);
        if (++slot == stages) { slot = 0; parity ^= 1u; }
    }
}

int fmt_of(DType dt) { return dt == DType::Q4_K_M ? 0 : dt == DType::Q5_K ? 1 : dt == DType::Q6_K ? 2 : -1; }

int g_num_sms = 0;
int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

constexpr size_t SMEM_CAP = 227 * 1024 - 256;    // static __shared__ red[] counts against the cap

template <int MASK, int WARPS>
void launch_kq(const KqParams& p, size_t smem, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        NT_CUDA_CHECK(cudaFuncSetAttribute(gemv_kq_kernel<MASK, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CAP));
        configured = true;
    }
    int grid = num_sms();
    const int need = (p.total_groups + WARPS - 1) / WARPS;
    if (need < grid) grid = need;
    launch_k(gemv_kq_kernel<MASK, WARPS>, dim3(grid), dim3(WARPS * 32), smem, s, p);
    count_launch();
}

template <int MASK>
void launch_fmt(KqParams& p, cudaStream_t s) {
    constexpr int SLOT = RG * BS * max_blk(MASK);
    const size_t xq_sz = ((size_t)3 * p.K + (size_t)(p.K / 32) * 4 + (size_t)(p.K / 16) * 4 + 127) & ~(size_t)127;
    const size_t budget = SMEM_CAP - xq_sz - 512;
    // CTA width: among the widths that afford a 2-deep ring, minimise rounds x warps (tail quantisation when
    // the SM is issue-bound), preferring wider CTAs on ties.
    const int sms = num_sms();
    int best_w = 0;
    long best_cost = 0;
    for (int w = 12; w >= 4; w--) {
        if (w == 11 || w == 9) continue;
        if ((size_t)w * 2 * (SLOT + 8) > budget) continue;
        const long rounds = (p.total_groups + (long)sms * w - 1) / ((long)sms * w);
        const long cost = rounds * w;
        if (!best_w || cost < best_cost) { best_w = w; best_cost = cost; }
    }
    { const char* f = getenv("NT_B200_GEMV_WARPS"); if (f && atoi(f) >= 4 && (size_t)atoi(f) * 2 * (SLOT + 8) <= budget) best_w = atoi(f); }
    NT_CHECK(best_w != 0, "gemv_kq: shared memory budget exceeded");
    int stages = (int)(budget / ((size_t)best_w * (SLOT + 8)));
    if (stages > 4) stages = 4;
    p.stages = stages;
    const size_t smem = xq_sz + (size_t)best_w * stages * (SLOT + 8) + 16;
    switch (best_w) {
        case 12: launch_kq<MASK, 12>(p, smem, s); break;
        case 10: launch_kq<MASK, 10>(p, smem, s); break;
        case 8: launch_kq<MASK, 8>(p, smem, s); break;
        case 7: launch_kq<MASK, 7>(p, smem, s); break;
        case 6: launch_kq<MASK, 6>(p, smem, s); break;
        case 5: launch_kq<MASK, 5>(p, smem, s); break;
        default: launch_kq<MASK, 4>(p, smem, s); break;
    }
}

}  // namespace

bool gemv_kq_supported(const GemvMat* mats, int n_mat, int K) {
    if (n_mat < 1 || n_mat > MAX_MATS || K % 256 != 0 || K <= 0) return false;
    if (3.375 * K > 120 * 1024) return false;   // xq must leave room for the rings
    for (int i = 0; i < n_mat; i++) {
        const int f = fmt_of(mats[i].dtype);
        if (f < 0 || mats[i].out <= 0) return false;
        const size_t pitch = mats[i].row_pitch ? mats[i].row_pitch : dtype_row_size(mats[i].dtype, K);
        if ((reinterpret_cast<uintptr_t>(mats[i].W) & 15) || (pitch & 15)) return false;
        // the last chunk of a row is copied in 16-byte units and must stay inside the row pitch
        const size_t blk = dtype_size(mats[i].dtype);
        const int NB = K / 256, NC = (NB + BS - 1) / BS, last = NB - (NC - 1) * BS;
        const size_t tail_end = (size_t)(NC - 1) * BS * blk + ((last * blk + 15) & ~(size_t)15);
        if (tail_end > pitch) return false;
    }
    return true;
}

void gemv_kq(const GemvMat* mats, int n_mat, int K, const GemvInput& in, GemvEpilogue ep, cudaStream_t s) {
    NT_CHECK(gemv_kq_supported(mats, n_mat, K), "gemv_kq: unsupported shape/dtype/alignment");
    NT_CHECK((in.xq != nullptr) != (in.x != nullptr), "gemv_kq: exactly one of xq / x must be given");
    if (ep == GEMV_SWIGLU)
        NT_CHECK(n_mat == 2 && mats[0].out == mats[1].out && mats[0].dtype == mats[1].dtype,
                 "gemv_kq: SWIGLU needs gate and up of equal rows and dtype");
    KqParams p{};
    p.K = K; p.NB = K / 256; p.NC = (p.NB + BS - 1) / BS;
    p.xq = static_cast<const int8_t*>(in.xq);
    p.x_f32 = in.x; p.norm_w = in.norm_w; p.eps = in.eps;
    p.epilogue = (int)ep;
    p.n_mat = n_mat;
    int total = 0, mask = 0;
    for (int i = 0; i < n_mat; i++) {
        KqMat& m = p.mat[i];
        m.W = static_cast<const uint8_t*>(mats[i].W);
        m.y = mats[i].y;
        m.out = mats[i].out;
        m.groups = (mats[i].out + RG - 1) / RG;
        m.fmt = fmt_of(mats[i].dtype);
        m.row_pitch = (long long)(mats[i].row_pitch ? mats[i].row_pitch : dtype_row_size(mats[i].dtype, K));
        total += m.groups;
        mask |= 1 << m.fmt;
    }
    if (ep == GEMV_SWIGLU) { p.n_seg = 2; p.total_groups = p.mat[0].groups; }
    else { p.n_seg = 1; p.total_groups = total; }
    switch (mask) {
        case 1: launch_fmt<1>(p, s); break;
        case 2: launch_fmt<2>(p, s); break;
        case 4: launch_fmt<4>(p, s); break;
        case 3: launch_fmt<3>(p, s); break;
        case 5: launch_fmt<5>(p, s); break;
        case 6: launch_fmt<6>(p, s); break;
        default: launch_fmt<7>(p, s); break;
    }
}

void gemv_kq(const GemvMat* mats, int n_mat, int K, const void* xq, GemvEpilogue ep, cudaStream_t s) {
    GemvInput in;
    in.xq = xq;
    gemv_kq(mats, n_mat, K, in, ep, s);
}

}}  // namespace nt::b200
