// gemv_kquant.cu — dequant-fused GEMV over GGUF Q4_K / Q5_K / Q6_K super-blocks for sm_100a.
//
// Replaces reference kernels K3/K4/K5 (src/cuda/gemm.cu:158-470) behind launch_gemv (gemm.cu:748-805).
// The reference walks one warp per row with per-lane byte loads, an F32 convert + 2 FMAs per weight and
// x re-read from shared memory for every row.  On B200 (7.3 TB/s reads, ~48 Q4_K weights/clk/SM) that
// instruction stream is ~2x over the issue budget, so this kernel is organised differently:
//
//   * chunk-stationary warps: a row is cut into chunks of 16 super-blocks (4096 weights); warp w of a CTA
//     owns chunk (w % NC) for the whole kernel, lane <-> one 128-weight half super-block.  The lane's slice of
//     the activation vector therefore never changes: it is loaded ONCE into registers (3 int8 planes x 128
//     elements = 96 registers) and the streaming loop touches shared memory only for weights;
//   * weights: each warp owns a private ring of TMA bulk copies (cp.async.bulk + mbarrier): one stage =
//     4 rows x 16 super-blocks (>= 2304 B per copy), filled by the warp's lane 0 and consumed by the same
//     warp, so HBM requests stay in flight across row-group boundaries with no CTA barrier on the data path;
//   * the NC warps that hold the chunks of one row-group meet once per row-group (a CTA barrier) and one of
//     them adds the NC partial sums in a fixed order (deterministic) and applies the epilogue.  All 148 CTAs
//     advance over the row-groups in lock-step rounds, so the tail is one row-group, not one warp-task;
//   * activations: block-scaled 3-term int8 (kernels_internal.h "xq"); math: IDP.4A on the 4/5/6-bit codes
//     (exact integer partial sums), one F32 scale per 16/32 weights, 6-bit scale unpack amortised over a
//     128-weight half super-block per lane.
// Up to 3 matrices of mixed K-quant formats share a launch (fused QKV of a Q4_K_M file; gate+up with SwiGLU).
// No tensor cores: the path is HBM-bound (BASELINE.json north_star).
#include "kernels_internal.h"
#include "ring.cuh"
#include "xquant.cuh"
#include "gemv_kq_device.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace nt { namespace b200 {

namespace {

constexpr int MAX_MATS = 3;
constexpr int MIN_WARPS = 4, MAX_WARPS = 12;


struct KqMat {
    const uint8_t* W;
    float* y;
    int out;
    int groups;        // ceil(out / RG)
    int fmt;           // 0 Q4_K, 1 Q5_K, 2 Q6_K, 3 Q8_0 (as groups of 8 blocks), 4 Q4_0 (likewise; opt-in, see fmt_of)
    long long row_pitch;
};
struct KqParams {
    KqMat mat[MAX_MATS];
    int n_mat;
    int K, NB, NC;             // elements, super-blocks per row, chunks per row (NC <= warps)
    const int8_t* xq;          // pre-quantised activations, or null:
    const float* x_f32;        //   F32 activations quantised in the prologue,
    const float* norm_w;       //   optionally as RMSNorm(x) * norm_w: the prologue quantises x * norm_w and sums x^2 in the same
    float eps;                 //   pass; the scalar rsqrt(mean(x^2) + eps) is applied to the results (exact up to round-off)
    int x_alias;               // the prologue's staging area aliases ring stages >= 1 (those are primed after the prologue)
    int total_groups;          // SWIGLU: groups of mat[0]
    int n_seg;                 // segments (matrices) per row-group: 2 for SWIGLU else 1
    int epilogue;
    int stages;                // ring depth per warp
    int gpc;                   // row-groups a CTA processes per round = warps / NC
    // Lock-step rounds over grid * gpc warp slots.  A partly filled last round would cost a whole round (70B gate/up: 7168
    // row-groups over 888 slots = 8.07 -> 9 rounds; o-projection 2.3 -> 3): its tail_groups row-groups are dealt out tail_nr (1 or 2)
    // rows at a time over all slots instead, so the tail costs tail_nr / 4 of a round.  tail_nr == 4: no split.
    int full_rounds, tail_groups, tail_nr;
    PeerOut peer;              // epilogue GEMV_PEER
};

// MASK: bit f set <=> matrices of format f may appear in this launch (mixed Q4_K_M projections share one launch).
template <int MASK, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 6 ? 2 : 1) gemv_kq_kernel(const __grid_constant__ KqParams p) {   // <= 6 warps: two CTAs may share an SM
    constexpr int SLOT = RG * BS * max_blk(MASK);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float red[32];
    __shared__ float partial[2][MAX_WARPS][2][RG];      // [buffer][warp][segment][row]
    const int K = p.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages, NC = p.NC, NB = p.NB, n_seg = p.n_seg, gpc = p.gpc;
    // smem carve-up: [ring stage 0: WARPS slots][stage 1: WARPS slots]...[mbarriers][x staging unless it aliases stages >= 1]
    uint8_t* ring = smem + (size_t)warp * SLOT;                             // this warp's slot of stage s: ring + s * WARPS * SLOT
    constexpr size_t STAGE_STRIDE = (size_t)WARPS * SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_STRIDE) + warp * stages;
    uint8_t* xs = p.x_alias ? smem + STAGE_STRIDE : smem + ((((size_t)WARPS * stages * (SLOT + 8)) + 127) & ~(size_t)127);

    const int chunk = warp % NC, gsub = warp / NC;        // this warp's chunk of every row, and its row-group slot
    const int n_rounds = p.full_rounds + (p.tail_groups > 0 ? 1 : 0);
    const int nbc = min(BS, NB - chunk * BS);              // super-blocks in this warp's chunk
    // what this warp slot holds in `round`: row-group g (global), the first row's offset inside it and the number of rows
    auto slot_of = [&](int round, int& g, int& row_sub, int& nr) -> bool {
        if (round < p.full_rounds || p.tail_nr >= RG) {
            g = (round * (int)gridDim.x + (int)blockIdx.x) * gpc + gsub; row_sub = 0; nr = RG;
            return g < p.total_groups;
        }
        nr = p.tail_nr;
        const int per = RG / nr, u = (int)blockIdx.x * gpc + gsub;
        g = p.full_rounds * (int)gridDim.x * gpc + u / per;
        row_sub = (u % per) * nr;
        return u < p.tail_groups * per;
    };
    auto locate = [&](int g, int seg, int& mi, int& gl) {  // matrix and local row-group of global group g
        if (n_seg == 2) { mi = seg; gl = g; return; }
        mi = 0;
        while (mi + 1 < p.n_mat && g >= p.mat[mi].groups) { g -= p.mat[mi].groups; mi++; }
        gl = g;
    };
    const int n_stages_total = n_rounds * n_seg;           // flattened stage index s = round * n_seg + seg
    // Producer cursor (used by lane 0): global row-group and segment of the next stage to fetch.  Kept incremental:
    // the TMA issue path runs on one lane but costs whole-warp issue slots (profiles/r01: 0.30 instr/weight before).
    int p_round = 0, p_seg = 0;
    auto issue_next = [&](int slot) {                      // lane 0: TMA copies of the next stage into ring slot
        uint64_t* bar = bars + slot;
        int g, row_sub, nr, mi = 0, gl = 0;
        bool have = slot_of(p_round, g, row_sub, nr);
        if (have) {
            locate(g, p_seg, mi, gl);
            have = gl * RG + row_sub < p.mat[mi].out;        // a slice of a ragged last group may hold no row
        }
        if (!have) {
            mbar_expect_tx(bar, 0);                          // nothing to fetch: just complete the phase
        } else {
            const KqMat& m = p.mat[mi];
            const int blkb = (MASK == 1) ? 144 : (MASK == 2) ? 176 : (MASK == 4) ? 210 : (MASK == 8) ? 272 : (MASK == 16) ? 144 : (m.fmt == 0 ? 144 : m.fmt == 1 ? 176 : 210);
            // copy size rounded up to 16 B (a 210-byte Q6_K tail may spill into row padding; checked on the host)
            const uint32_t bytes = ((uint32_t)(nbc * blkb) + 15u) & ~15u;
            mbar_expect_tx(bar, bytes * nr);
            uint8_t* dst = ring + (size_t)slot * STAGE_STRIDE;
            const int row0 = gl * RG + row_sub;
            const uint8_t* src = m.W + (long long)chunk * (BS * blkb) + (long long)row0 * m.row_pitch;
            if (nr == RG && row0 + RG <= m.out) {
#pragma unroll
                for (int r = 0; r < RG; r++) bulk_g2s(dst + r * (BS * blkb), src + r * m.row_pitch, bytes, bar);
            } else {                                         // tail stage / ragged last group: re-read the last valid row
                for (int r = 0; r < nr; r++)
                    bulk_g2s(dst + r * (BS * blkb), src + (long long)min(r, m.out - 1 - row0) * m.row_pitch, bytes, bar);
            }
        }
        if (++p_seg == n_seg) { p_seg = 0; p_round++; }
    };

    if (lane == 0) {
        for (int s = 0; s < stages; s++) mbar_init(bars + s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    int issued = 0;
    // Weights do not depend on the previous kernel: start streaming before touching x.  When the prologue's staging area
    // aliases the ring, only stage 0 is primed here and the rest follows once x sits in registers.
    const int prime_now = p.x_alias ? 1 : stages;
    if (lane == 0) {
        for (; issued < prime_now && issued < n_stages_total; issued++) issue_next(issued);
    }
    pdl_wait();   // no-op unless launched with programmatic stream serialization

    // ---- this lane's activation slice -> registers (once) ----
    const int blk = lane >> 1, h = lane & 1;
    const uint32_t hb = (uint32_t)((chunk * BS + min(blk, nbc - 1)) * 2 + h);   // global half-block index (clamped for idle lanes)
    XRegs X;
    float rms_inv = 1.0f;
    if (p.x_f32) {
        // Fused prologue (the stateless launch_gemv path and the decode chain): F32 vector -> xq form in shared memory, one pass.
        // With norm_w the quantiser sees x * norm_w and the pass also sums x^2; block-scaled quantisation is scale invariant,
        // so multiplying the RESULTS by rsqrt(mean(x^2) + eps) equals RMSNorm-then-GEMV up to round-off (rmsnorm.cu:17-70).
        float* xscale = reinterpret_cast<float*>(xs + 3 * (size_t)K);
        float* xsum16 = xscale + K / 32;
        // One thread <-> 8 consecutive elements (4 threads share a 32-element block: 2 shuffles for its absmax, 1 for the
        // 16-element sums), all loads of the pass issued up front: the whole CTA is busy and the pass costs one L2 round trip.
        // Same representation as quantize_lane32 (x ~= scale * (q1 * 16384 + q2 * 128 + q3)), with x / s taken as x * (127 / amax).
        constexpr int MAXV = 3;                             // thread-iterations held in registers at once (K <= 24 * threads)
        const int nvec = K / 8, nthr = WARPS * 32;
        float ss = 0.f;
        for (int base = 0; base < nvec; base += MAXV * nthr) {
            float4 hv[MAXV][2], wv[MAXV][2];
#pragma unroll
            for (int j = 0; j < MAXV; j++) {
                const int t = base + j * nthr + (int)threadIdx.x;
                if (t < nvec) {
                    hv[j][0] = __ldcg(reinterpret_cast<const float4*>(p.x_f32) + 2 * t);
                    hv[j][1] = __ldcg(reinterpret_cast<const float4*>(p.x_f32) + 2 * t + 1);
                    if (p.norm_w) {
                        wv[j][0] = __ldg(reinterpret_cast<const float4*>(p.norm_w) + 2 * t);
                        wv[j][1] = __ldg(reinterpret_cast<const float4*>(p.norm_w) + 2 * t + 1);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < MAXV; j++) {
                const int t = base + j * nthr + (int)threadIdx.x;
                if (t >= nvec) break;                       // nvec % 32 == 0: whole warps leave together
                float x[8] = {hv[j][0].x, hv[j][0].y, hv[j][0].z, hv[j][0].w, hv[j][1].x, hv[j][1].y, hv[j][1].z, hv[j][1].w};
                if (p.norm_w) {
                    const float w[8] = {wv[j][0].x, wv[j][0].y, wv[j][0].z, wv[j][0].w, wv[j][1].x, wv[j][1].y, wv[j][1].z, wv[j][1].w};
#pragma unroll
                    for (int i = 0; i < 8; i++) { ss = fmaf(x[i], x[i], ss); x[i] = __fmul_rn(x[i], w[i]); }
                }
                float amax = 0.f, s8 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) { amax = fmaxf(amax, fabsf(x[i])); s8 = __fadd_rn(s8, x[i]); }
                amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, 1));
                amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, 2));
                const float s16 = __fadd_rn(s8, __shfl_xor_sync(0xFFFFFFFFu, s8, 1));
                const float sc = __fdiv_rn(amax, 127.0f);
                const float inv = (amax > 0.f) ? __fdiv_rn(127.0f, amax) : 0.f;
                uint32_t w1[2] = {0, 0}, w2[2] = {0, 0}, w3[2] = {0, 0};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float tq = __fmul_rn(x[i], inv);
                    const float f1 = rintf(tq);
                    const float r1 = __fmul_rn(__fsub_rn(tq, f1), 128.0f);
                    const float f2 = rintf(r1);
                    const float r2 = __fmul_rn(__fsub_rn(r1, f2), 128.0f);
                    const float f3 = rintf(r2);
                    w1[i >> 2] |= ((uint32_t)(int)f1 & 0xFFu) << (8 * (i & 3));
                    w2[i >> 2] |= ((uint32_t)(int)f2 & 0xFFu) << (8 * (i & 3));
                    w3[i >> 2] |= ((uint32_t)(int)f3 & 0xFFu) << (8 * (i & 3));
                }
                const uint32_t se = xq_swizzle((uint32_t)t * 8u);          // 8 bytes stay inside one swizzled 16-byte column
                *reinterpret_cast<uint2*>(xs + se) = make_uint2(w1[0], w1[1]);
                *reinterpret_cast<uint2*>(xs + K + se) = make_uint2(w2[0], w2[1]);
                *reinterpret_cast<uint2*>(xs + 2 * (size_t)K + se) = make_uint2(w3[0], w3[1]);
                if ((threadIdx.x & 3) == 0) xscale[t >> 2] = __fmul_rn(sc, 1.0f / 16384.0f);
                if ((threadIdx.x & 1) == 0) xsum16[t >> 1] = s16;
            }
        }
        if (p.norm_w) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
            if (lane == 0) red[warp] = ss;
        }
        __syncthreads();
        if (p.norm_w) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < WARPS; i++) t += red[i];    // fixed order: the same value in every warp and CTA
            rms_inv = rsqrtf(t / K + p.eps);
        }
        {
            const uint32_t sw = (hb & 7u) << 4;
#pragma unroll
            for (int pl = 0; pl < 3; pl++)
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int4 v = *reinterpret_cast<const int4*>(xs + (size_t)pl * K + ((hb * 128u + 16u * i) ^ sw));
                    X.x[pl][4 * i] = v.x; X.x[pl][4 * i + 1] = v.y; X.x[pl][4 * i + 2] = v.z; X.x[pl][4 * i + 3] = v.w;
                }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) X.sx[j] = xscale[hb * 4 + j];
#pragma unroll
        for (int j = 0; j < 8; j++) X.s16[j] = xsum16[hb * 8 + j];
        if (p.x_alias) {
            __syncthreads();                                 // every lane holds its slice: the staging area becomes ring again
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes before async-proxy (TMA) writes
                for (; issued < stages && issued < n_stages_total; issued++) issue_next(issued);
            }
        }
    } else {
        // pre-quantised xq in global memory (planes stored with the 16-byte-column swizzle of xquant.cuh)
        const int8_t* xq = p.xq;
        const uint32_t sw = (hb & 7u) << 4;
#pragma unroll
        for (int pl = 0; pl < 3; pl++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(xq + (size_t)pl * K + ((hb * 128u + 16u * i) ^ sw)));
                X.x[pl][4 * i] = v.x; X.x[pl][4 * i + 1] = v.y; X.x[pl][4 * i + 2] = v.z; X.x[pl][4 * i + 3] = v.w;
            }
        const float* fs = reinterpret_cast<const float*>(xq + 3 * (size_t)K);
        const float4 sv = __ldg(reinterpret_cast<const float4*>(fs + hb * 4));
        X.sx[0] = sv.x; X.sx[1] = sv.y; X.sx[2] = sv.z; X.sx[3] = sv.w;
        const float4 u0 = __ldg(reinterpret_cast<const float4*>(fs + K / 32 + hb * 8));
        const float4 u1 = __ldg(reinterpret_cast<const float4*>(fs + K / 32 + hb * 8 + 4));
        X.s16[0] = u0.x; X.s16[1] = u0.y; X.s16[2] = u0.z; X.s16[3] = u0.w;
        X.s16[4] = u1.x; X.s16[5] = u1.y; X.s16[6] = u1.z; X.s16[7] = u1.w;
    }
    unsigned peer_seq = 0, peer_parity = 0;
    if (p.epilogue == GEMV_PEER) { peer_seq = __ldcg(p.peer.seq) + 1u; peer_parity = peer_seq & 1u; }
    pdl_launch_dependents();

    int slot = 0;
    uint32_t parity = 0;
    for (int round = 0; round < n_rounds; round++) {
        int g, row_sub, nr;
        const bool live = slot_of(round, g, row_sub, nr);
        float res[2] = {0.f, 0.f};
        for (int seg = 0; seg < n_seg; seg++) {
            float acc[RG] = {0.f, 0.f, 0.f, 0.f};
            mbar_wait(bars + slot, parity);
            if (live && blk < nbc) {
                int mi, gl;
                locate(g, seg, mi, gl);
                const uint8_t* slot_base = ring + (size_t)slot * STAGE_STRIDE;
                const int fmt = (MASK == 1) ? 0 : (MASK == 2) ? 1 : (MASK == 4) ? 2 : (MASK == 8) ? 3 : (MASK == 16) ? 4 : p.mat[mi].fmt;
                if (nr == RG) {
                    if ((MASK & 1) && fmt == 0) process_stage<0>(slot_base, blk, h, X, acc);
                    if ((MASK & 2) && fmt == 1) process_stage<1>(slot_base, blk, h, X, acc);
                    if ((MASK & 4) && fmt == 2) process_stage<2>(slot_base, blk, h, X, acc);
                    if ((MASK & 8) && fmt == 3) process_stage<3>(slot_base, blk, h, X, acc);
                    if ((MASK & 16) && fmt == 4) process_stage<4>(slot_base, blk, h, X, acc);
                } else {                                     // tail stage of 1 or 2 rows: one row at a time
                    const int rowp = BS * ((fmt == 0 || fmt == 4) ? 144 : fmt == 1 ? 176 : fmt == 2 ? 210 : 272);
#pragma unroll 1
                    for (int r = 0; r < nr; r++) {
                        float a1[RG] = {0.f, 0.f, 0.f, 0.f};
                        const uint8_t* rb = slot_base + r * rowp;
                        if ((MASK & 1) && fmt == 0) process_stage<0, 1>(rb, blk, h, X, a1);
                        if ((MASK & 2) && fmt == 1) process_stage<1, 1>(rb, blk, h, X, a1);
                        if ((MASK & 4) && fmt == 2) process_stage<2, 1>(rb, blk, h, X, a1);
                        if ((MASK & 8) && fmt == 3) process_stage<3, 1>(rb, blk, h, X, a1);
                        if ((MASK & 16) && fmt == 4) process_stage<4, 1>(rb, blk, h, X, a1);
                        acc[0] = (r == 0) ? a1[0] : acc[0];       // no dynamic indexing: acc stays in registers
                        acc[1] = (r == 1) ? a1[0] : acc[1];
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && issued < n_stages_total) issue_next(slot);
            issued++;
            if (++slot == stages) { slot = 0; parity ^= 1u; }
            res[seg] = reduce4(acc, lane);
        }
        // ---- combine the NC chunk partials of each row-group (fixed order => deterministic) ----
        const int buf = round & 1;
        if (live && (lane & 7) == 0) {
            const int r = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
            partial[buf][warp][0][r] = res[0];
            partial[buf][warp][1][r] = res[1];
        }
        __syncthreads();
        if (live && chunk == 0 && lane < nr) {
            float v0 = 0.f, v1 = 0.f;
            for (int c = 0; c < NC; c++) {
                v0 += partial[buf][gsub * NC + c][0][lane];
                if (n_seg == 2) v1 += partial[buf][gsub * NC + c][1][lane];
            }
            v0 *= rms_inv; v1 *= rms_inv;                   // 1.0 unless the prologue normalised (RMSNorm's scalar factor)
            int mi, gl;
            locate(g, 0, mi, gl);
            const KqMat& m = p.mat[mi];
            const int row = gl * RG + row_sub + lane;
            if (row < m.out) {
                if (p.epilogue == GEMV_SWIGLU) {
                    // silu(g) * u with the reference's fast-math expression (gemm.cu:713-725)
                    m.y[row] = __fdividef(v0, 1.0f + __expf(-v0)) * v1;
                } else if (p.epilogue == GEMV_ADD) {
                    m.y[row] += v0;
                } else if (p.epilogue == GEMV_PEER) {
                    // this rank's partial row -> slot [parity][rank] on every rank: one 64-bit NVLink store {sequence : value} per peer
                    // (lanes 0..3 write 32 contiguous bytes); the value is its own arrival flag
                    const size_t off = ((size_t)peer_parity * p.peer.size + p.peer.rank) * (size_t)p.peer.hidden + (size_t)row;
                    const unsigned long long pkt = ((unsigned long long)peer_seq << 32) | (unsigned long long)__float_as_uint(v0);
                    for (int r = 0; r < p.peer.size; r++)
                        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.peer.slots[r] + off), "l"(pkt) : "memory");
                } else {
                    m.y[row] = v0;
                }
            }
        }
    }
}

// Q4_0 (reference K1, gemm.cu:32-86) runs on this path by default since its first hardware run (round 2: tests/test_q4_0_tma_gpu.py
// against the oracle, <= 2e-5); NT_B200_Q4_0_TMA=0 sends it back to the generic kernel (gemv_generic.cu) for A/B comparisons.
bool q4_0_tma_enabled() {
    static const bool on = [] { const char* e = getenv("NT_B200_Q4_0_TMA"); return !(e && e[0] == '0' && e[1] == 0); }();
    return on;
}
int fmt_of(DType dt) {
    return dt == DType::Q4_K_M ? 0 : dt == DType::Q5_K ? 1 : dt == DType::Q6_K ? 2 : dt == DType::Q8_0 ? 3
         : (dt == DType::Q4_0 && q4_0_tma_enabled()) ? 4 : -1;
}

int g_num_sms = 0;
int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

constexpr size_t SMEM_CAP = 227 * 1024 - 1024;    // static __shared__ (red, partial) counts against the cap

bool quarter_enabled() {          // NT_B200_GEMV_QB=1: opt-in (measured slower than the half-block kernel, see below)
    static const bool on = [] { const char* e = getenv("NT_B200_GEMV_QB"); return e && *e && !(e[0] == '0' && e[1] == 0); }();
    return on;
}

bool quarter_split_enabled() {    // NT_B200_GEMV_QB_SPLIT=0: launches that mix Q6_K with Q4_K/Q5_K stay one half-block launch
    static const bool on = [] { const char* e = getenv("NT_B200_GEMV_QB_SPLIT"); return !(e && e[0] == '0' && e[1] == 0); }();
    return on;
}

bool tail_split_enabled() {       // NT_B200_TAIL_SPLIT=0 keeps whole 4-row stages in the last round (A/B aid)
    static const bool on = [] { const char* e = getenv("NT_B200_TAIL_SPLIT"); return !(e && e[0] == '0' && e[1] == 0); }();
    return on;
}

template <int MASK, int WARPS>
void launch_kq(const KqParams& p, size_t smem, cudaStream_t s) {
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(gemv_kq_kernel<MASK, WARPS>, (int)((int)SMEM_CAP), configured);
    int grid = num_sms();
    KqParams q = p;
    const int slots = grid * q.gpc;
    q.full_rounds = q.total_groups / slots;
    q.tail_groups = q.total_groups % slots;
    q.tail_nr = RG;
    if (q.tail_groups > 0 && tail_split_enabled()) {          // also when the whole launch is less than one round (narrow shards)
        if (4 * q.tail_groups <= slots) q.tail_nr = 1;
        else if (2 * q.tail_groups <= slots) q.tail_nr = 2;
    }
    if (q.full_rounds == 0) {                                  // fewer units than warp slots: no more CTAs than there is work
        const int units = q.tail_groups * (RG / q.tail_nr);
        const int need = (units + q.gpc - 1) / q.gpc;
        if (need < grid) grid = need;
    }
    launch_k(gemv_kq_kernel<MASK, WARPS>, dim3(grid), dim3(WARPS * 32), smem, s, q);
    count_launch();
}

// Warps per CTA: the largest multiple of NC (chunks per row) in [MIN_WARPS, MAX_WARPS] that affords `min_stages` ring
// stages.  Measured on the 70B shapes (profiles/r01_gemv_width_scan.txt): 12 warps beat 8 and 10 on the fused gate+up
// GEMV (55.6 vs 59.8 us) even though the last round is less full, so occupancy wins over tail balance.
int pick_warps(int NC, int total_groups, size_t slot, size_t budget, int min_stages) {
    (void)total_groups;
    const char* force = getenv("NT_B200_GEMV_WARPS");               // tuning aid
    int best = 0;
    for (int w = NC; w <= MAX_WARPS; w += NC) {
        if (w < MIN_WARPS) continue;
        if ((size_t)w * min_stages * (slot + 8) > budget) break;
        if (force && atoi(force) == w) return w;
        best = w;
    }
    return best;
}

template <int MASK>
void launch_fmt(KqParams& p, cudaStream_t s) {
    constexpr int SLOT = RG * BS * max_blk(MASK);
    const size_t xq_sz = p.x_f32 ? (((size_t)3 * p.K + (size_t)(p.K / 32) * 4 + (size_t)(p.K / 16) * 4 + 127) & ~(size_t)127) + 128 : 0;
    // F32 input: first try to stage x over ring stages >= 1 (full ring depth and CTA width, the ring is primed in two steps);
    // vectors too long for that get their own area behind a smaller ring.
    int w = 0, stages = 0;
    p.x_alias = 0;
    if (p.x_f32) {
        const size_t budget = SMEM_CAP - 256;
        w = pick_warps(p.NC, p.total_groups, SLOT, budget, 2);
        if (w) {
            stages = (int)(budget / ((size_t)w * (SLOT + 8)));
            if (stages > 4) stages = 4;
            if (xq_sz <= (size_t)w * (stages - 1) * SLOT) p.x_alias = 1; else w = 0;
        }
    }
    if (!w) {
        const size_t budget = SMEM_CAP - xq_sz - 256;
        w = pick_warps(p.NC, p.total_groups, SLOT, budget, 2);
        if (!w) w = pick_warps(p.NC, p.total_groups, SLOT, budget, 1);   // long rows with in-kernel quantisation: single-stage ring
        NT_CHECK(w != 0, "gemv_kq: shared memory budget exceeded");
        stages = (int)(budget / ((size_t)w * (SLOT + 8)));
        if (stages > 4) stages = 4;
    }
    static const int stage_cap = [] { const char* e = getenv("NT_B200_GEMV_STAGES"); return e ? atoi(e) : 0; }();   // tuning aid: ring depth cap
    if (stage_cap >= 2 && stages > stage_cap && (!p.x_alias || ((size_t)w * (stage_cap - 1) * SLOT >= xq_sz))) stages = stage_cap;
    p.gpc = w / p.NC;
    p.stages = stages;
    const size_t smem = (size_t)w * stages * (SLOT + 8) + 128 + (p.x_alias ? 0 : xq_sz);
    switch (w) {
        case 12: launch_kq<MASK, 12>(p, smem, s); break;
        case 11: launch_kq<MASK, 11>(p, smem, s); break;
        case 10: launch_kq<MASK, 10>(p, smem, s); break;
        case 9: launch_kq<MASK, 9>(p, smem, s); break;
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
    const int NB = K / 256, NC = (NB + BS - 1) / BS;
    if (NC > MAX_WARPS) return false;              // one warp per chunk of a row (K <= 49152)
    for (int i = 0; i < n_mat; i++) {
        const int f = fmt_of(mats[i].dtype);
        if (f < 0 || mats[i].out <= 0) return false;
        const size_t pitch = mats[i].row_pitch ? mats[i].row_pitch : dtype_row_size(mats[i].dtype, K);
        if ((reinterpret_cast<uintptr_t>(mats[i].W) & 15) || (pitch & 15)) return false;
        // the last chunk of a row is copied in 16-byte units and must stay inside the row pitch
        const size_t blk = dtype_row_size(mats[i].dtype, 256);     // bytes per 256 weights
        const int last = NB - (NC - 1) * BS;
        const size_t tail_end = (size_t)(NC - 1) * BS * blk + ((last * blk + 15) & ~(size_t)15);
        if (tail_end > pitch) return false;
    }
    return true;
}

void gemv_kq(const GemvMat* mats, int n_mat, int K, const GemvInput& in, GemvEpilogue ep, cudaStream_t s) {
    NT_CHECK(gemv_kq_supported(mats, n_mat, K), "gemv_kq: unsupported shape/dtype/alignment");
    NT_CHECK(ep != GEMV_PEER || (in.peer && n_mat == 1), "gemv_kq: GEMV_PEER needs GemvInput::peer and a single matrix");
    // Opt-in A/B (NT_B200_GEMV_QB=1): the quarter-block kernel (gemv_kquant_q.cu: 16 warps per SM, 14 on the 28672-wide down
    // projection); launches that mix Q6_K with another format are split per matrix for it.  Measured on the 70B shapes
    // (profiles/r02_qb_layer_ncu_summary.txt): correct (all kernel / model tests) but slower — gate+up 62.0 us vs 53.2, Q6_K down
    // 49.8 vs 45.0, step 84.0 vs 86.3 tok/s on the same box: the per-lane header / scale work is paid per 64 instead of per 128
    // weights (+25 % instructions), which more than eats the fourth warp per scheduler.  The half-block kernel stays the default.
    if (quarter_enabled()) {
        int mask = 0;
        for (int i = 0; i < n_mat; i++) mask |= 1 << fmt_of(mats[i].dtype);
        const bool one_family = mask == 1 || mask == 2 || mask == 3 || mask == 4 || mask == 8 || mask == 16;
        if (one_family) {
            if (gemv_kq_quarter(mats, n_mat, K, in, ep, in.peer, s)) return;
        } else if (ep != GEMV_SWIGLU && ep != GEMV_PEER && quarter_split_enabled()) {
            // q/k (Q4_K) + v (Q6_K) of a Q4_K_M file: the Q4_K/Q5_K matrices in one launch, each Q6_K matrix in its own
            GemvMat fam[MAX_MATS];
            int nf = 0;
            for (int i = 0; i < n_mat; i++) if (mats[i].dtype != DType::Q6_K) fam[nf++] = mats[i];
            int fmask = 0;
            for (int i = 0; i < nf; i++) fmask |= 1 << fmt_of(fam[i].dtype);
            if (nf > 0 && (fmask == 1 || fmask == 2 || fmask == 3)) {
                bool ok = gemv_kq_quarter(fam, nf, K, in, ep, nullptr, s);
                for (int i = 0; i < n_mat && ok; i++)
                    if (mats[i].dtype == DType::Q6_K) ok = gemv_kq_quarter(&mats[i], 1, K, in, ep, nullptr, s);
                NT_CHECK(ok, "gemv_kq: the quarter-block kernel took part of a split launch only");
                return;
            }
        }
    }
    NT_CHECK((in.xq != nullptr) != (in.x != nullptr), "gemv_kq: exactly one of xq / x must be given");
    if (ep == GEMV_SWIGLU)
        NT_CHECK(n_mat == 2 && mats[0].out == mats[1].out, "gemv_kq: SWIGLU needs gate and up of equal rows");
    int mask = 0;
    for (int i = 0; i < n_mat; i++) mask |= 1 << fmt_of(mats[i].dtype);
    const bool plain = mask == 1 || mask == 2 || mask == 4 || mask == 8 || mask == 3 || mask == 5 || mask == 16;
    if (!plain && ep != GEMV_SWIGLU) {                         // rare mixes: one launch per matrix
        for (int i = 0; i < n_mat; i++) gemv_kq(&mats[i], 1, K, in, ep, s);
        return;
    }
    NT_CHECK(plain, "gemv_kq: SWIGLU over this format mix is not instantiated");
    KqParams p{};
    p.K = K; p.NB = K / 256; p.NC = (p.NB + BS - 1) / BS;
    p.xq = static_cast<const int8_t*>(in.xq);
    p.x_f32 = in.x; p.norm_w = in.norm_w; p.eps = in.eps;
    if (ep == GEMV_PEER) p.peer = *in.peer;
    p.epilogue = (int)ep;
    p.n_mat = n_mat;
    int total = 0;
    for (int i = 0; i < n_mat; i++) {
        KqMat& m = p.mat[i];
        m.W = static_cast<const uint8_t*>(mats[i].W);
        m.y = mats[i].y;
        m.out = mats[i].out;
        m.groups = (mats[i].out + RG - 1) / RG;
        m.fmt = fmt_of(mats[i].dtype);
        m.row_pitch = (long long)(mats[i].row_pitch ? mats[i].row_pitch : dtype_row_size(mats[i].dtype, K));
        total += m.groups;
    }
    if (ep == GEMV_SWIGLU) { p.n_seg = 2; p.total_groups = p.mat[0].groups; }
    else { p.n_seg = 1; p.total_groups = total; }
    switch (mask) {
        case 1: launch_fmt<1>(p, s); break;
        case 2: launch_fmt<2>(p, s); break;
        case 4: launch_fmt<4>(p, s); break;
        case 3: launch_fmt<3>(p, s); break;
        case 8: launch_fmt<8>(p, s); break;
        case 16: launch_fmt<16>(p, s); break;
        default: launch_fmt<5>(p, s); break;
    }
}

void gemv_kq_peer(const GemvMat& mat, int K, const GemvInput& in, const PeerOut& peer, cudaStream_t s) {
    NT_CHECK(mat.out == peer.hidden && peer.size >= 2 && peer.size <= PeerOut::kMaxTP, "gemv_kq_peer: rows must equal the exchanged vector length");
    GemvInput ip = in;
    ip.peer = &peer;
    gemv_kq(&mat, 1, K, ip, GEMV_PEER, s);
}

void gemv_kq(const GemvMat* mats, int n_mat, int K, const void* xq, GemvEpilogue ep, cudaStream_t s) {
    GemvInput in;
    in.xq = xq;
    gemv_kq(mats, n_mat, K, in, ep, s);
}

}}  // namespace nt::b200
