// decode_megakernel.cu — see decode_mega.h for the design and the (un)tested status.
//
// Every phase is a transplant of a kernel of the decode graph, so that with one rank and split_fixed set to the graph
// path's split count the two paths can be compared bit for bit:
//   MPH_NORM_XQ  <- rmsnorm_xq_kernel      (elementwise.cu; reference K12 rmsnorm.cu:17-70)
//   MPH_QUANT    <- quantize_x_kernel      (elementwise.cu)
//   MPH_GEMV     <- gemv_kq_kernel         (gemv_kquant.cu; reference K2-K5 gemm.cu:96-470)
//   MPH_ATTN     <- rope_kv_decode_kernel + decode_kernel (elementwise.cu, attention.cu; reference K14/K15/K16
//                   rotary.cu:16-62, attention.cu:316-342, attention.cu:108-202)
//   MPH_COMBINE  <- decode_combine_kernel  (attention.cu)
// Compiled with --use_fast_math like those files so the transcendental expressions lower to the same instructions.
// Beyond the transplant (all optional, MegaFuse bits in decode_mega.h): last-arriver fusions of the activation quantiser, the
// split combine and (one rank) the residual add + next norm into their producers; MPH_REDUCE_XQ, the distributed form of the
// norm phase with the 1/rms factor applied by the consumer (always used under tensor parallelism); weight prefetch across the
// attention phase.  File layout: memory-model helpers and barriers, the GEMV phase (its schedule lives in decode_mega_sched.h,
// shared with the host-side plan builder and schedule replay in decode_mega_plan.cu), norm / quantise / reduce phases, attention,
// the kernel and, last, the CUDA-runtime-facing DecodeMega class.  tests/cusim compiles this file for the CPU emulator.
#include "decode_mega_sched.h"
#include "../ring.cuh"
#include "../xquant.cuh"
#include <cuda_fp16.h>
#include <cfloat>
#include <cstring>
#include <algorithm>

namespace nt { namespace b200 {

namespace {

#ifdef NT_CUSIM
#define NT_NOINLINE __attribute__((noinline))
#else
#define NT_NOINLINE __noinline__
#endif

constexpr int NTHREADS = MEGA_NTHREADS;
constexpr int AW = MEGA_ATTN_WARPS;
constexpr int KU = 4;                                                        // K / V rows in flight per warp in the attention phase

// ---- memory-model helpers -----------------------------------------------------------------------------------------
#ifdef NT_CUSIM   // CPU emulation (tests/cusim): C++ atomics; the acquire loads yield so that spinning threads let others run
inline unsigned ld_acquire_gpu(const unsigned* p) { cusim::yield("spin (grid barrier word)"); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned ld_relaxed_sys(const unsigned* p) { cusim::yield("spin (peer flag)"); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned ld_relaxed_gpu(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void red_release_gpu_add(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline void red_relaxed_gpu_add(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline void st_release_gpu(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void st_relaxed_sys(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void red_relaxed_sys_add(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline unsigned long long global_timer_ns() { return (unsigned long long)(cusim::now_s() * 1e9); }
inline void fence_proxy_async_smem() {}
inline void prefetch_l2(const void*) {}
inline unsigned long long sm_clock() { return (unsigned long long)(cusim::now_s() * 1e9); }
inline void bulk_prefetch_l2(const void*, unsigned) {}
#else
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_sys_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long sm_clock() { return (unsigned long long)clock64(); }
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {    // src and bytes: multiples of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

#endif

// Spin until (int)(*p - target) >= 0.  SYS: the word is written by another GPU and polled with relaxed system-scope loads (the
// caller issues one system fence after all of its peers have arrived); otherwise acquire loads at GPU scope.  A time-out (or
// an abort raised elsewhere) sets the abort word and returns: every later barrier then falls through and the host reports it.
template <bool SYS>
__device__ NT_NOINLINE void spin_until(const unsigned* p, unsigned target, unsigned* abort_word, unsigned long long timeout_ns, unsigned tag) {
    if (ld_relaxed_gpu(abort_word)) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned n = 0;
    for (;;) {
        const unsigned v = SYS ? ld_relaxed_sys(p) : ld_acquire_gpu(p);
        if ((int)(v - target) >= 0) return;
        if ((++n & 255u) == 0) {
            if (ld_relaxed_gpu(abort_word)) return;
            if (global_timer_ns() - t0 > timeout_ns) {
                if (atomicExch(abort_word, 1u + (SYS ? 1u : 0u)) == 0u) {    // first to give up: leave a note for the host
                    abort_word[1] = tag;                                      // barrier index of this launch
                    abort_word[2] = blockIdx.x;
                    abort_word[3] = target;
                    abort_word[4] = SYS ? ld_relaxed_sys(p) : ld_relaxed_gpu(p);
                }
                return;
            }
        }
    }
}

struct SyncState {
    unsigned bar_idx;      // grid barriers passed in this launch
    unsigned xchg_idx;     // tensor-parallel exchanges passed in this launch
    unsigned xchg_base;    // exchange sequence number at launch (monotonic across launches)
};

// Grid barrier (+ tensor-parallel flag exchange when kind == MBAR_EXCHANGE and tp_size > 1).
__device__ void mega_barrier(const MegaParams& P, int kind, SyncState& st) {
    if (kind == MBAR_NONE) return;
    st.bar_idx++;
    const bool xchg = (kind == MBAR_EXCHANGE) && P.tp_size > 1;
    if (xchg) st.xchg_idx++;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* counter = P.sync;
        unsigned* go = P.sync + 32;
        unsigned* abort_word = P.sync + 64;
        const unsigned target = st.bar_idx * gridDim.x;
        // One fence per hop.  Arrive: release at GPU scope, or — when this CTA has just pushed rows into peer memory — a
        // system fence followed by a relaxed arrive (the system fence subsumes the GPU-scope release).
        if (xchg) { __threadfence_system(); red_relaxed_gpu_add(counter, 1u); }
        else red_release_gpu_add(counter, 1u);
        if (xchg && P.xchg_direct) {
            // Experiment (off by default; KNOWN GAP): with tp_size >= 3 a rank that has already passed exchange n can post
            // its exchange-(n+1) arrivals into a slower rank's counter while another rank's exchange-n arrival is still in flight
            // (relaxed reds to different destinations are unordered), so the counter can reach n * tp * grid early.  Needs one
            // counter per exchange parity (or per source rank) before it may be used; the stand-alone exchange of
            // engine/peer_xchg.cu avoids the issue altogether by making every value its own flag.
            // No master hop.  Every CTA adds 1 to every rank's arrival counter (remote atomics over NVLink, posted) and
            // waits until its own rank's counter has seen all tp_size * grid CTAs of this exchange.  The counter is never reset:
            // the target grows with the exchange sequence number (a peer can be at most one exchange ahead).
            for (int r = 0; r < P.tp_size; r++) red_relaxed_sys_add(P.flags[r] + 32 * P.tp_size, 1u);
            const unsigned all = (st.xchg_base + st.xchg_idx) * (unsigned)P.tp_size * gridDim.x;
            spin_until<true>(P.flags[P.tp_rank] + 32 * P.tp_size, all, abort_word, P.timeout_ns, st.bar_idx);
            __threadfence_system();                                                 // acquire side: the peers' rows are visible
        } else if (xchg) {
            const unsigned seq = st.xchg_base + st.xchg_idx;
            if (blockIdx.x == 0) {
                spin_until<false>(counter, target, abort_word, P.timeout_ns, st.bar_idx);       // every local CTA has pushed its rows
                __threadfence_system();                                             // ... before the flags become visible
                for (int r = 0; r < P.tp_size; r++)
                    if (r != P.tp_rank) st_relaxed_sys(P.flags[r] + 32 * P.tp_rank, seq);      // posted NVLink writes
                for (int r = 0; r < P.tp_size; r++)
                    if (r != P.tp_rank) spin_until<true>(P.flags[P.tp_rank] + 32 * r, seq, abort_word, P.timeout_ns, st.bar_idx);   // local polls
                __threadfence_system();                                             // acquire side: the peers' rows are visible
                st_release_gpu(go, st.bar_idx);
            } else {
                spin_until<false>(go, st.bar_idx, abort_word, P.timeout_ns, st.bar_idx);
            }
        } else {
            spin_until<false>(counter, target, abort_word, P.timeout_ns, st.bar_idx);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ---- shared-memory state -------------------------------------------------------------------------------------------
struct Shared {
    MegaPhase ph[4];                                 // descriptor ring: phase i lives in ph[i & 3], prefetched two phases ahead
    uint64_t bars[MEGA_WARPS * MEGA_MAX_STAGES];     // one mbarrier per (warp, ring stage)
    float partial[2][MEGA_WARPS][2][RG];             // [buffer][warp][segment][row] chunk partial sums
    float red[32];
    float s_max[8], s_sum[8];
    unsigned bcast;                                  // CTA-wide broadcast of a last-arriver ticket
};
static_assert(sizeof(Shared) <= MEGA_STATIC_SMEM, "static shared memory budget");

__device__ __forceinline__ void load_phase(MegaPhase* dst, const MegaPhase* src) {
    constexpr int WORDS = (int)(sizeof(MegaPhase) / 4);
    if ((int)threadIdx.x < WORDS) reinterpret_cast<int*>(dst)[threadIdx.x] = reinterpret_cast<const int*>(src)[threadIdx.x];
}

// Lane 0 of an active warp: TMA copies of the producer's next stage into ring slot `slot`; every lane advances the cursor.
__device__ __forceinline__ void issue_next(const MegaPhase& d, Producer& pr, uint8_t* ring, uint64_t* bars, int slot, int chunk,
                                           int gsub, int nbc, int lane) {
    if (lane == 0) {
        uint64_t* bar = bars + slot;
        const StageRef sr = stage_ref(d, (int)gridDim.x, (int)blockIdx.x, gsub, pr.round, pr.seg, chunk, nbc);
        if (sr.empty) {
            mbar_expect_tx(bar, 0);                          // nothing to fetch: just complete the phase
        } else {
            const MegaMat& m = d.mat[sr.mi];
            mbar_expect_tx(bar, sr.bytes * sr.nrows);
            uint8_t* dst = ring + (size_t)slot * d.slot_bytes;
            const uint8_t* src = m.W + sr.src_off;
            const int last = m.out - 1 - sr.row0;            // rows past the end of a ragged group re-read the last valid row
            for (int r = 0; r < sr.nrows; r++)
                bulk_g2s(dst + r * (BS * sr.blkb), src + (long long)min(r, last) * m.pitch, sr.bytes, bar);
        }
    }
    producer_advance(d, pr);
}

// Start streaming the weights of GEMV phase `d` (already in shared memory): fill this warp's ring.
__device__ __forceinline__ void prime_rings(const MegaPhase& d, Producer& pr, uint8_t* smem, uint64_t* bars_all, int warp, int lane) {
    pr.issued = 0; pr.round = 0; pr.seg = 0;
    if (warp >= d.warps) return;
    const int chunk = warp % d.NC, gsub = warp / d.NC;
    const int nbc = min(BS, d.NB - chunk * BS);
    uint8_t* ring = smem + (size_t)warp * d.stages * d.slot_bytes;
    uint64_t* bars = bars_all + warp * MEGA_MAX_STAGES;
    const int n_total = d.n_rounds * d.n_seg;
    if (lane == 0) fence_proxy_async_smem();                 // the ring area may have been written by generic-proxy stores
    for (; pr.issued < d.stages && pr.issued < n_total; pr.issued++) issue_next(d, pr, ring, bars, pr.issued, chunk, gsub, nbc, lane);
}

// ---- GEMV phase: consumer ------------------------------------------------------------------------------------------------
__device__ void gemv_phase(const MegaParams& P, Shared& S, const MegaPhase& d, uint8_t* smem, Producer& pr, uint32_t& parity_bits,
                           int warp, int lane) {
    const int n_rounds = d.n_rounds;
    float rms_inv = 1.0f;
    if (d.ssq_in) {
        // MEGA_FUSE_NORM consumer: the activations were quantised as h * norm_w; the missing 1/rms factor is a scalar of the
        // whole vector and is applied to the results.  Every CTA adds the per-block sums of squares in the same order.
        float part = 0.f;
        for (int b = threadIdx.x; b < P.hidden / 32; b += NTHREADS) part += __ldcg(d.ssq_in + b);
        part = warp_sum(part);
        if (lane == 0) S.red[warp] = part;
        __syncthreads();
        float t = (lane < MEGA_WARPS) ? S.red[lane] : 0.f;
        t = warp_sum(t);
        rms_inv = rsqrtf(t / P.hidden + P.eps);
    }
    if (warp >= d.warps) {                                   // idle warp of this phase: keep the CTA barriers balanced
        for (int round = 0; round < n_rounds; round++) __syncthreads();
        return;
    }
    const int K = d.K, NC = d.NC, n_seg = d.n_seg, stages = d.stages;
    const int chunk = warp % NC, gsub = warp / NC;
    const int nbc = min(BS, d.NB - chunk * BS);
    uint8_t* ring = smem + (size_t)warp * stages * d.slot_bytes;
    uint64_t* bars = S.bars + warp * MEGA_MAX_STAGES;
    const int n_stages_total = n_rounds * n_seg;

    // ---- this lane's activation slice -> registers (once per phase); xq was written earlier in this kernel: L2 loads ----
    const int blk = lane >> 1, h = lane & 1;
    const uint32_t hb = (uint32_t)((chunk * BS + min(blk, nbc - 1)) * 2 + h);
    XRegs X;
    {
        const int8_t* xq = d.xq;
        const uint32_t sw = (hb & 7u) << 4;
#pragma unroll
        for (int pl = 0; pl < 3; pl++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int4 v = __ldcg(reinterpret_cast<const int4*>(xq + (size_t)pl * K + ((hb * 128u + 16u * i) ^ sw)));
                X.x[pl][4 * i] = v.x; X.x[pl][4 * i + 1] = v.y; X.x[pl][4 * i + 2] = v.z; X.x[pl][4 * i + 3] = v.w;
            }
        const float* fs = reinterpret_cast<const float*>(xq + 3 * (size_t)K);
        const float4 sv = __ldcg(reinterpret_cast<const float4*>(fs + hb * 4));
        X.sx[0] = sv.x; X.sx[1] = sv.y; X.sx[2] = sv.z; X.sx[3] = sv.w;
        const float4 u0 = __ldcg(reinterpret_cast<const float4*>(fs + K / 32 + hb * 8));
        const float4 u1 = __ldcg(reinterpret_cast<const float4*>(fs + K / 32 + hb * 8 + 4));
        X.s16[0] = u0.x; X.s16[1] = u0.y; X.s16[2] = u0.z; X.s16[3] = u0.w;
        X.s16[4] = u1.x; X.s16[5] = u1.y; X.s16[6] = u1.z; X.s16[7] = u1.w;
    }

    if ((d.fuse & MEGA_L2_PREFETCH) && warp == d.warps - 1 && lane == 0) {
        // experiment (wide tensor parallelism, where a layer's shard fits the L2): this CTA's 1/grid slice of the next GEMV
        // phase's matrices -> L2, fire and forget, while this phase computes
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const unsigned long long total = d.pf_bytes[i];
            if (!total) continue;
            const unsigned long long per = ((total + gridDim.x - 1) / gridDim.x + 15ull) & ~15ull;
            unsigned long long off = per * blockIdx.x;
            const unsigned long long end = min(total, off + per);
            for (; off < end; off += 32768ull) bulk_prefetch_l2(d.pf_ptr[i] + off, (unsigned)min(32768ull, end - off));
        }
    }
    int slot = 0;
    for (int round = 0; round < n_rounds; round++) {
        // this slot's work in the round (segment 1, when there is one, is the same rows of the second matrix)
        const StageRef sr = stage_ref(d, (int)gridDim.x, (int)blockIdx.x, gsub, round, 0, chunk, nbc);
        const bool live = !sr.empty;
        float res[2] = {0.f, 0.f};
        for (int seg = 0; seg < n_seg; seg++) {
            float acc[RG] = {0.f, 0.f, 0.f, 0.f};
            mbar_wait(bars + slot, (parity_bits >> slot) & 1u);
            parity_bits ^= 1u << slot;
            if (live && blk < nbc) {
                const uint8_t* slot_base = ring + (size_t)slot * d.slot_bytes;
                const int fmt = d.mat[n_seg == 2 ? seg : sr.mi].fmt;
                if (sr.nrows == RG) {
                    if (fmt == 0) process_stage<0>(slot_base, blk, h, X, acc);
                    else if (fmt == 1) process_stage<1>(slot_base, blk, h, X, acc);
                    else if (fmt == 2) process_stage<2>(slot_base, blk, h, X, acc);
                    else if (fmt == 3) process_stage<3>(slot_base, blk, h, X, acc);
                    else process_stage<4>(slot_base, blk, h, X, acc);
                } else {                                     // tail stage of 1 or 2 rows: one row at a time
                    const int rowp = BS * blk_bytes(fmt);
#pragma unroll 1
                    for (int r = 0; r < sr.nrows; r++) {
                        float a1[RG] = {0.f, 0.f, 0.f, 0.f};
                        const uint8_t* rb = slot_base + r * rowp;
                        if (fmt == 0) process_stage<0, 1>(rb, blk, h, X, a1);
                        else if (fmt == 1) process_stage<1, 1>(rb, blk, h, X, a1);
                        else if (fmt == 2) process_stage<2, 1>(rb, blk, h, X, a1);
                        else if (fmt == 3) process_stage<3, 1>(rb, blk, h, X, a1);
                        else process_stage<4, 1>(rb, blk, h, X, a1);
                        acc[r] = a1[0];
                    }
                }
            }
            __syncwarp();
            if (pr.issued < n_stages_total) issue_next(d, pr, ring, bars, slot, chunk, gsub, nbc, lane);
            pr.issued++;
            if (++slot == stages) slot = 0;
            res[seg] = reduce4(acc, lane);
        }
        // ---- combine the NC chunk partials of each row-group (fixed order => deterministic) ----
        const int buf = round & 1;
        if (live && (lane & 7) == 0) {
            const int r = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
            S.partial[buf][warp][0][r] = res[0];
            S.partial[buf][warp][1][r] = res[1];
        }
        __syncthreads();
        if (live && chunk == 0) {                            // warp-uniform
            const MegaMat& m = d.mat[sr.mi];
            const int n_valid = min(sr.nrows, m.out - sr.row0);       // rows of this stage that exist (ragged last group)
            if (lane < n_valid) {
                float v0 = 0.f, v1 = 0.f;
                for (int c = 0; c < NC; c++) {
                    v0 += S.partial[buf][gsub * NC + c][0][lane];
                    if (n_seg == 2) v1 += S.partial[buf][gsub * NC + c][1][lane];
                }
                v0 *= rms_inv; v1 *= rms_inv;                // 1.0 unless the input came from a fused norm
                const int row = sr.row0 + lane;
                if (d.epilogue == MEP_SWIGLU) {
                    m.y[row] = __fdividef(v0, 1.0f + __expf(-v0)) * v1;           // reference gemm.cu:713-725
                } else if (d.epilogue == MEP_SLOT) {
                    // partial result of this rank -> slot [parity][this rank] on every tensor-parallel peer (NVLink stores)
                    const size_t off = ((size_t)d.slot_parity * P.tp_size + P.tp_rank) * (size_t)P.hidden + (size_t)row;
                    for (int r = 0; r < P.tp_size; r++) P.slots[r][off] = v0;
                } else {
                    m.y[row] = v0;
                }
            }
            const bool fq = d.epilogue == MEP_SWIGLU && (d.fuse & MEGA_FUSE_QUANT);
            const bool fn = d.epilogue == MEP_SLOT && (d.fuse & MEGA_FUSE_NORM);
            if (fq || fn) {
                // Last arriver of a 32-row block (tickets count rows: stages of the tail round hold fewer than RG).  Writers fence
                // before the ticket, the last arriver fences before it reads the block back.
                __threadfence();
                __syncwarp();
                const int block = sr.row0 >> 5;
                const int rows_in_block = min(32, m.out - block * 32);
                unsigned prev = 0;
                if (lane == 0) prev = atomicAdd(d.cnt + block, (unsigned)n_valid);
                prev = __shfl_sync(0xFFFFFFFFu, prev, 0);
                if ((int)prev + n_valid == rows_in_block) {
                    __threadfence();
                    const int e = block * 32 + lane;
                    if (fq) {
                        // the SwiGLU output block -> xq (same quantize_block32 as the stand-alone phase => same bytes): the down
                        // projection needs no separate quantise phase and barrier
                        const float v = (e < d.n) ? __ldcg(d.x + e) : 0.f;
                        quantize_block32(v, block, lane, d.xq_out, d.n);
                    } else {
                        // single rank: the slot rows ARE the full projection.  Residual add, the block's sum of squares, and the
                        // next norm's quantiser input h * w (1/rms is applied by the consumer)
                        const float* sl = P.slots[P.tp_rank] + (size_t)d.slot_parity * P.tp_size * (size_t)P.hidden;
                        const float hval = __ldcg(d.hid_in + e) + __ldcg(sl + e);
                        d.hid_out[e] = hval;
                        const float ss = warp_sum(hval * hval);
                        if (lane == 0) d.ssq_out[block] = ss;
                        quantize_block32(hval * d.norm_w[e], block, lane, d.xq_out, P.hidden);
                    }
                    if (lane == 0) d.cnt[block] = 0;         // ready for the next layer
                }
            }
        }
    }
    if (d.epilogue == MEP_SLOT && P.tp_size > 1 && chunk == 0 && lane < RG) __threadfence_system();   // the lanes that stored to peer memory
}

// ---- norm + quantise (distributed over the first hidden/256 CTAs, as rmsnorm_xq_kernel) -----------------------------------
__device__ __forceinline__ float residual_at(const MegaParams& P, const MegaPhase& d, int i) {
    float v = __ldcg(d.hid_in + i);
    if (d.pending_parity >= 0) {
        // sum of the ranks' partial rows in rank order; the loads are issued together (one L2 round trip, not tp_size)
        const float* sl = P.slots[P.tp_rank] + (size_t)d.pending_parity * P.tp_size * (size_t)P.hidden + i;
        float sv[MEGA_MAX_TP];
#pragma unroll
        for (int r = 0; r < MEGA_MAX_TP; r++) sv[r] = (r < P.tp_size) ? __ldcg(sl + (size_t)r * P.hidden) : 0.f;
        float t = sv[0];
#pragma unroll
        for (int r = 1; r < MEGA_MAX_TP; r++) if (r < P.tp_size) t += sv[r];
        v += t;
    }
    return v;
}

// Distributed form of the norm phase (MEGA_DEFER_RMS): one warp per 32-element block anywhere in the grid, a single L2 round
// trip, no CTA-wide reduction.  The 1/rms factor is left to the consumer (ssq_in), exactly as for MEGA_FUSE_NORM.
__device__ void reduce_xq_phase(const MegaParams& P, const MegaPhase& d, int warp, int lane) {
    const int nblk = P.hidden / 32;
    for (int b = (int)blockIdx.x * MEGA_WARPS + warp; b < nblk; b += (int)gridDim.x * MEGA_WARPS) {
        const int e = b * 32 + lane;
        const float hval = residual_at(P, d, e);
        if (d.hid_out) d.hid_out[e] = hval;
        const float ss = warp_sum(hval * hval);
        if (lane == 0) d.ssq_out[b] = ss;
        quantize_block32(hval * d.norm_w[e], b, lane, d.xq_out, P.hidden);
    }
}

__device__ void norm_xq_phase(const MegaParams& P, Shared& S, const MegaPhase& d, int warp, int lane) {
    const int hidden = P.hidden;
    if ((int)blockIdx.x * 256 >= hidden) return;             // CTA-uniform
    float ss = 0.f;
    if (threadIdx.x < 256)
        for (int i = threadIdx.x; i < hidden; i += 256) { float v = residual_at(P, d, i); ss += v * v; }
    ss = warp_sum(ss);
    __syncthreads();
    if (lane == 0 && warp < 8) S.red[warp] = ss;
    __syncthreads();
    float t = (lane < 8) ? S.red[lane] : 0.f;
    t = warp_sum(t);
    const float rms_inv = rsqrtf(t / hidden + P.eps);
    const int blk = (int)blockIdx.x * 8 + warp;
    if (warp >= 8 || blk * 32 >= hidden) return;
    const int e = blk * 32 + lane;
    const float hval = residual_at(P, d, e);
    if (d.hid_out) d.hid_out[e] = hval;
    const float v = hval * rms_inv * d.norm_w[e];
    quantize_block32(v, blk, lane, d.xq_out, hidden);
}

__device__ void quant_phase(const MegaPhase& d, int warp, int lane) {
    const int nblk = d.n / 32;
    for (int b = (int)blockIdx.x * MEGA_WARPS + warp; b < nblk; b += (int)gridDim.x * MEGA_WARPS)
        quantize_block32(__ldcg(d.x + b * 32 + lane), b, lane, d.xq_out, d.n);
}

// ---- attention ------------------------------------------------------------------------------------------------------------------
template <int N> struct HalfVec;
template <> struct HalfVec<2> { using T = uint32_t; };
template <> struct HalfVec<4> { using T = uint2; };
template <> struct HalfVec<8> { using T = uint4; };

template <int DPL>
__device__ __forceinline__ void load_row(const __half* p, float (&f)[DPL]) {
    typename HalfVec<DPL>::T raw = *reinterpret_cast<const typename HalfVec<DPL>::T*>(p);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < DPL / 2; i++) { float2 t = __half22float2(h2[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

template <int GC>
__device__ __forceinline__ float transpose_reduce(float (&v)[GC], int lane) {
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
    float r = v[0];
#pragma unroll
    for (int off = 16 / GC; off > 0; off >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, off);
    return r;
}

struct SplitRule { int split_len, used; };
__device__ __forceinline__ SplitRule split_rule(const MegaParams& P, int ctx, int n_groups) {
    SplitRule r;
    if (P.split_fixed > 0) {
        r.split_len = max((ctx + P.split_fixed - 1) / P.split_fixed, MEGA_COMPAT_MIN_SPLIT);   // the graph path's rule (attention.cu DYN_MIN_SPLIT)
    } else {
        const int cap = max(1, (int)gridDim.x / n_groups);
        r.split_len = max(P.min_split, (ctx + cap - 1) / cap);
        r.split_len = min(r.split_len, P.max_split);
    }
    r.used = (ctx + r.split_len - 1) / r.split_len;
    return r;
}

__device__ __forceinline__ void combine_head(const MegaParams& P, int h, int used);

// One unit: GC query heads that share KV head `kv_head`, keys [k_begin, k_end) of a context of pos + 1 tokens.
// RoPE of the unit's queries and of the token's own key (reference rotary.cu:16-62 expression), the F16 rounding of the
// cache write (attention.cu:316-342) and, in the unit that owns the last split, the cache write itself happen here, so
// no grid barrier is needed between the q/k/v projection and attention.  Scores / softmax / P.V as attend_group.
template <int DPL, int GC>
__device__ void attend_unit(const MegaParams& P, Shared& S, const MegaPhase& d, float* smf, int head0, int kv_head, int split,
                            int k_begin, int k_end, int pos, bool write_cache, int grp, int used) {
    constexpr int HD = DPL * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_keys = k_end - k_begin;
    float* sc = smf;                                         // [GC][max_split]
    float* red = sc + (size_t)GC * P.max_split;              // [AW][GC][HD]
    float* q_s = red + (size_t)AW * GC * HD;                 // [GC][HD] rotated queries
    __half* cur_k = reinterpret_cast<__half*>(q_s + GC * HD);   // [HD] the token's own key / value, rounded like the cache
    __half* cur_v = cur_k + HD;
    const int n_kv = P.nkv;
    const __half* kc = static_cast<const __half*>(d.kc);
    const __half* vc = static_cast<const __half*>(d.vc);

    // ---- RoPE (q of the GC heads, k of the KV head) + this token's cache row ----
    {
        const int head_dim = P.hd;                           // run-time value: same division as rope_kv_decode_kernel
        const int half_dim = head_dim / 2;
        const int total_q = GC * half_dim;
        for (int idx = threadIdx.x; idx < total_q + half_dim; idx += NTHREADS) {
            const bool is_key = idx >= total_q;
            const int li = is_key ? idx - total_q : idx;
            const int pair = li % half_dim, head = li / half_dim;
            float freq = 1.0f / powf(P.theta, (2.0f * pair) / head_dim);
            float angle = pos * freq * P.freq_scale;
            float c = cosf(angle), sn = sinf(angle);
            const float* src = is_key ? P.k + (size_t)kv_head * head_dim : P.q + (size_t)(head0 + head) * head_dim;
            float x0 = __ldcg(src + pair), x1 = __ldcg(src + pair + half_dim);
            float r0 = x0 * c - x1 * sn, r1 = x1 * c + x0 * sn;
            if (!is_key) {
                q_s[head * HD + pair] = r0;
                q_s[head * HD + pair + half_dim] = r1;
            } else {
                const __half k0 = __float2half(r0), k1 = __float2half(r1);
                const __half v0 = __float2half(__ldcg(P.v + (size_t)kv_head * head_dim + pair));
                const __half v1 = __float2half(__ldcg(P.v + (size_t)kv_head * head_dim + pair + half_dim));
                cur_k[pair] = k0; cur_k[pair + half_dim] = k1;
                cur_v[pair] = v0; cur_v[pair + half_dim] = v1;
                if (write_cache && pos < P.max_seq) {
                    const size_t row = (size_t)pos * n_kv * head_dim + (size_t)kv_head * head_dim;
                    __half* kw = static_cast<__half*>(d.kc);
                    __half* vw = static_cast<__half*>(d.vc);
                    kw[row + pair] = k0; kw[row + pair + half_dim] = k1;
                    vw[row + pair] = v0; vw[row + pair + half_dim] = v1;
                }
            }
        }
    }
    __syncthreads();

    const size_t row_stride = (size_t)n_kv * HD;
    const __half* kbase = kc + (size_t)kv_head * HD + (size_t)lane * DPL;
    const __half* vbase = vc + (size_t)kv_head * HD + (size_t)lane * DPL;
    if (warp < AW) {
        float qr[GC][DPL];
#pragma unroll
        for (int g = 0; g < GC; g++)
#pragma unroll
            for (int i = 0; i < DPL; i++) qr[g][i] = q_s[g * HD + lane * DPL + i];
        // ---- phase 1: scores.  KU key rows are fetched before any of them is used: one memory round trip per KU keys ----
        for (int p0 = warp; p0 < n_keys; p0 += AW * KU) {
            float kf[KU][DPL];
#pragma unroll
            for (int u = 0; u < KU; u++) {
                const int p = p0 + u * AW;
                if (p < n_keys) {
                    if (k_begin + p == pos) load_row<DPL>(cur_k + lane * DPL, kf[u]);
                    else {
                        load_row<DPL>(kbase + (size_t)(k_begin + p) * row_stride, kf[u]);
                        // the matching V row is needed two CTA barriers from now: start pulling it from HBM into L2
                        if ((lane & (32 / DPL - 1)) == 0) prefetch_l2(vbase + (size_t)(k_begin + p) * row_stride);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < KU; u++) {
                const int p = p0 + u * AW;
                if (p < n_keys) {                            // warp-uniform
                    float part[GC];
#pragma unroll
                    for (int g = 0; g < GC; g++) {
                        float a = 0.f;
#pragma unroll
                        for (int i = 0; i < DPL; i++) a = fmaf(qr[g][i], kf[u][i], a);
                        part[g] = a;
                    }
                    float tot = transpose_reduce<GC>(part, lane);
                    if ((lane & (32 / GC - 1)) == 0) sc[(size_t)(lane / (32 / GC)) * P.max_split + p] = tot * P.attn_scale;
                }
            }
        }
    }
    __syncthreads();
    // ---- phase 2: per-head max / exp / sum (warp g <-> head g) ----
    if (warp < AW) {
        for (int g = warp; g < GC; g += AW) {
            float* s = sc + (size_t)g * P.max_split;
            float mx = -FLT_MAX;
            for (int p = lane; p < n_keys; p += 32) mx = fmaxf(mx, s[p]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
            float sum = 0.f;
            for (int p = lane; p < n_keys; p += 32) { float e = expf(s[p] - mx); s[p] = e; sum += e; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
            if (lane == 0) { S.s_max[g] = mx; S.s_sum[g] = sum; }
        }
    }
    __syncthreads();
    // ---- phase 3: P.V (warp <-> keys, lane <-> DPL dims) ----
    if (warp < AW) {
        float acc[GC][DPL];
#pragma unroll
        for (int g = 0; g < GC; g++)
#pragma unroll
            for (int i = 0; i < DPL; i++) acc[g][i] = 0.f;
        for (int p0 = warp; p0 < n_keys; p0 += AW * KU) {    // same key order per warp as the one-at-a-time loop
            float vf[KU][DPL];
#pragma unroll
            for (int u = 0; u < KU; u++) {
                const int p = p0 + u * AW;
                if (p < n_keys) {
                    if (k_begin + p == pos) load_row<DPL>(cur_v + lane * DPL, vf[u]);
                    else load_row<DPL>(vbase + (size_t)(k_begin + p) * row_stride, vf[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < KU; u++) {
                const int p = p0 + u * AW;
                if (p < n_keys) {
#pragma unroll
                    for (int g = 0; g < GC; g++) {
                        float w = sc[(size_t)g * P.max_split + p];
#pragma unroll
                        for (int i = 0; i < DPL; i++) acc[g][i] = fmaf(w, vf[u][i], acc[g][i]);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < GC; g++)
#pragma unroll
            for (int i = 0; i < DPL; i++) red[((size_t)warp * GC + g) * HD + lane * DPL + i] = acc[g][i];
    }
    __syncthreads();
    if ((d.fuse & MEGA_FUSE_COMBINE) && used == 1) {
        // The only split of its head group: nothing to merge.  Normalise and quantise straight from shared memory — the values
        // the combine would produce (o = t * e^0, l = sum * e^0), without the round trip through the scratch buffer.
        if (threadIdx.x < AW * 32) {
            for (int idx = threadIdx.x; idx < GC * HD; idx += AW * 32) {      // 32 consecutive idx = one warp, one head
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < AW; w++) t += red[(size_t)w * GC * HD + idx];
                const int g = idx / HD, dd = idx - g * HD;
                const float l = S.s_sum[g];
                const float v = t * ((l > 0.f) ? 1.0f / l : 0.f);
                const int e = (head0 + g) * HD + dd;
                P.attn_out[e] = v;
                quantize_block32(v, e >> 5, lane, P.xq_a, P.nh * HD);
            }
        }
        __syncthreads();                                     // shared memory is reused by the CTA's next unit
        return;
    }
    // ---- unnormalised partials of this split -> scratch [head][split][HD], [head][split][2] ----
    const int NS = P.n_splits_max;
    float* ml = P.attn_scratch + (size_t)P.nh * NS * HD;
    if (threadIdx.x < AW * 32) {
        for (int idx = threadIdx.x; idx < GC * HD; idx += AW * 32) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < AW; w++) t += red[(size_t)w * GC * HD + idx];
            const int g = idx / HD, dd = idx - g * HD;
            P.attn_scratch[((size_t)(head0 + g) * NS + split) * HD + dd] = t;
        }
        if (threadIdx.x < 2 * GC) {
            const int g = threadIdx.x >> 1;
            ml[((size_t)(head0 + g) * NS + split) * 2 + (threadIdx.x & 1)] = (threadIdx.x & 1) ? S.s_sum[g] : S.s_max[g];
        }
    }
    if (d.fuse & MEGA_FUSE_COMBINE) {
        // Last-arriver combine: the unit that completes the last split of this head group merges the group's splits and
        // emits their xq (same arithmetic as the combine phase), so no separate combine phase and barrier.
        __threadfence();                                     // this unit's partials are visible before its ticket
        __syncthreads();
        if (threadIdx.x == 0) S.bcast = atomicAdd(d.cnt + grp, 1u);
        __syncthreads();
        if ((int)S.bcast == used - 1) {                      // CTA-uniform
            __threadfence();
            if (threadIdx.x < 128)
                for (int g = 0; g < GC; g++) combine_head(P, head0 + g, used);
            if (threadIdx.x == 0) d.cnt[grp] = 0;            // ready for the next layer
        }
    }
    __syncthreads();                                         // shared memory is reused by the CTA's next unit
}

template <int DPL, int GC>
__device__ void attn_phase(const MegaParams& P, Shared& S, const MegaPhase& d, uint8_t* smem) {
    const int pos = P.step[1], ctx = pos + 1;
    const int n_groups = P.nh / GC, ratio = P.nh / P.nkv;
    const SplitRule sr = split_rule(P, ctx, n_groups);
    const int units = n_groups * sr.used;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u / n_groups, grp = u - split * n_groups;
        const int head0 = grp * GC, kv_head = head0 / ratio;
        const int k_begin = split * sr.split_len, k_end = min(ctx, k_begin + sr.split_len);
        // the unit that holds the token's own position writes the cache row; one writer per KV head
        const bool write_cache = (k_end == ctx) && (head0 % ratio == 0);
        attend_unit<DPL, GC>(P, S, d, reinterpret_cast<float*>(smem + P.attn_smem_off), head0, kv_head, split, k_begin, k_end, pos, write_cache, grp, sr.used);
    }
}

// Before the barrier that precedes an attention phase: the K/V cache rows of the units this CTA is about to process do not
// depend on the token being computed, so start pulling them from HBM into L2 while the grid waits (hints only).
__device__ void attn_prefetch(const MegaParams& P, const MegaPhase& next) {
    const int pos = P.step[1], ctx = pos + 1;
    const int n_groups = P.nh / P.gc, ratio = P.nh / P.nkv;
    const SplitRule sr = split_rule(P, ctx, n_groups);
    const int units = n_groups * sr.used;
    const size_t row_bytes = (size_t)P.nkv * P.hd * sizeof(__half), head_bytes = (size_t)P.hd * sizeof(__half);
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u / n_groups, grp = u - split * n_groups;
        const int kv_head = (grp * P.gc) / ratio;
        const int k_begin = split * sr.split_len, k_end = min(pos, k_begin + sr.split_len);      // the token's own row is not in the cache yet
        const char* kb = static_cast<const char*>(next.kc) + (size_t)kv_head * head_bytes;
        const char* vb = static_cast<const char*>(next.vc) + (size_t)kv_head * head_bytes;
        const int lines = (int)((head_bytes + 127) / 128);
        for (int t = threadIdx.x; t < (k_end - k_begin) * lines; t += NTHREADS) {
            const size_t off = (size_t)(k_begin + t / lines) * row_bytes + (size_t)(t % lines) * 128;
            prefetch_l2(kb + off);
            prefetch_l2(vb + off);
        }
    }
}

// Merge the split partials of head h and emit its slice of the o-projection's xq (decode_combine_kernel with xq_out).
// Called by threads 0..127 of a CTA (hd % 32 == 0: whole warps stay together in the quantiser).
__device__ __forceinline__ void combine_head(const MegaParams& P, int h, int used) {
    constexpr int CU = 8;                                    // splits whose partials are in flight together
    const int hd = P.hd, n_heads = P.nh, NS = P.n_splits_max;
    const float* ml = P.attn_scratch + (size_t)n_heads * NS * hd + (size_t)h * NS * 2;
    float m = -FLT_MAX;
    for (int i0 = 0; i0 < used; i0 += CU) {
        float a[CU];
#pragma unroll
        for (int j = 0; j < CU; j++) a[j] = (i0 + j < used) ? __ldcg(ml + 2 * (i0 + j)) : -FLT_MAX;
#pragma unroll
        for (int j = 0; j < CU; j++) m = fmaxf(m, a[j]);
    }
    float l = 0.f;
    for (int i0 = 0; i0 < used; i0 += CU) {
        float a[CU], b[CU];
#pragma unroll
        for (int j = 0; j < CU; j++) {
            a[j] = (i0 + j < used) ? __ldcg(ml + 2 * (i0 + j)) : 0.f;
            b[j] = (i0 + j < used) ? __ldcg(ml + 2 * (i0 + j) + 1) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < CU; j++) if (i0 + j < used) l += b[j] * expf(a[j] - m);
    }
    const float inv = (l > 0.f) ? 1.0f / l : 0.f;
    for (int dd = threadIdx.x; dd < hd; dd += 128) {
        float o = 0.f;
        for (int i0 = 0; i0 < used; i0 += CU) {
            float a[CU], b[CU];
#pragma unroll
            for (int j = 0; j < CU; j++) {
                a[j] = (i0 + j < used) ? __ldcg(ml + 2 * (i0 + j)) : 0.f;
                b[j] = (i0 + j < used) ? __ldcg(P.attn_scratch + ((size_t)h * NS + i0 + j) * hd + dd) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < CU; j++) if (i0 + j < used) o += b[j] * expf(a[j] - m);
        }
        const float v = o * inv;
        P.attn_out[(size_t)h * hd + dd] = v;
        const int K = n_heads * hd, e = h * hd + dd, lane = threadIdx.x & 31;
        quantize_block32(v, e >> 5, lane, P.xq_a, K);
    }
}

__device__ void combine_phase(const MegaParams& P) {
    const int ctx = P.step[1] + 1;
    const SplitRule sr = split_rule(P, ctx, P.nh / P.gc);
    if (threadIdx.x >= 128) return;
    for (int h = blockIdx.x; h < P.nh; h += gridDim.x) combine_head(P, h, sr.used);
}

// ---- the kernel ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) decode_step_kernel(const __grid_constant__ MegaParams P) {
#ifdef NT_CUSIM
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(cusim::g_cta->dyn_smem.data()) + 127) & ~(uintptr_t)127);
#else
    extern __shared__ __align__(128) uint8_t smem[];
#endif
    __shared__ Shared S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    SyncState st;
    st.bar_idx = 0; st.xchg_idx = 0;
    st.xchg_base = (P.tp_size > 1) ? P.sync[96] : 0u;        // written by the previous launch (stream order)
    uint32_t parity_bits = 0;                                // mbarrier phase parity of this warp's ring stages
    Producer pr;
    pr.issued = 0; pr.round = 0; pr.seg = 0;

    if (lane == 0) {
        for (int s = 0; s < MEGA_MAX_STAGES; s++) mbar_init(S.bars + warp * MEGA_MAX_STAGES + s, 1);
        mbar_fence_init();
    }
    // Descriptor ring: phase i is read from S.ph[i & 3].  Descriptors 0 and 1 are loaded here, descriptor i + 2 at the start
    // of phase i (its slot held phase i - 2, which every thread has left), so no phase waits for its own descriptor and the
    // priming step at the end of phase i (target i + 1 or i + 2, host-checked) finds its descriptor in shared memory.
    load_phase(&S.ph[0], P.phases);
    if (P.n_phases > 1) load_phase(&S.ph[1], P.phases + 1);
    __syncthreads();
    // Weights do not depend on anything computed in this kernel: start streaming the first GEMV's rows right away.
    if (P.first_gemv >= 0 && P.first_gemv < P.n_phases) prime_rings(S.ph[P.first_gemv & 3], pr, smem, S.bars, warp, lane);

    const bool tracing = P.trace != nullptr && blockIdx.x < MEGA_TRACE_CTAS && threadIdx.x == 0;
    // per traced CTA: [phase][start, work done, barrier passed] in SM clock ticks (%globaltimer only ticks every ~1 us), then
    // [clock at start, clock at end, globaltimer at start, globaltimer at end] to convert ticks to time
    unsigned long long* trace = tracing ? P.trace + (size_t)blockIdx.x * P.trace_stride : nullptr;
    unsigned long long* trace_cal = tracing ? trace + P.trace_stride - 4 : nullptr;
    if (tracing) { trace_cal[0] = sm_clock(); trace_cal[2] = global_timer_ns(); }
    for (int i = 0; i < P.n_phases; i++) {
        if (i + 2 < P.n_phases) load_phase(&S.ph[(i + 2) & 3], P.phases + i + 2);
        const MegaPhase& d = S.ph[i & 3];
        const int kind = d.kind;
        if (tracing) trace[3 * i] = sm_clock();
        if (kind == MPH_GEMV) {
            gemv_phase(P, S, d, smem, pr, parity_bits, warp, lane);
        } else if (kind == MPH_NORM_XQ) {
            norm_xq_phase(P, S, d, warp, lane);
        } else if (kind == MPH_QUANT) {
            quant_phase(d, warp, lane);
        } else if (kind == MPH_REDUCE_XQ) {
            reduce_xq_phase(P, d, warp, lane);
        } else if (kind == MPH_ATTN) {
            if (P.hd == 128 && P.gc == 8) attn_phase<4, 8>(P, S, d, smem);
            else if (P.hd == 128 && P.gc == 4) attn_phase<4, 4>(P, S, d, smem);
            else attn_phase<2, 4>(P, S, d, smem);            // hd 64, 4 query heads per KV head (host-checked)
        } else {
            combine_phase(P);
        }
        const int prime = d.prime;
        if (prime >= 0 && prime < P.n_phases) {
            __syncthreads();                                 // every warp is done with the ring area; descriptor i + 2 is visible
            prime_rings(S.ph[prime & 3], pr, smem, S.bars, warp, lane);
        }
        if (i + 1 < P.n_phases && S.ph[(i + 1) & 3].kind == MPH_ATTN) attn_prefetch(P, S.ph[(i + 1) & 3]);
        if (tracing) trace[3 * i + 1] = sm_clock();
        mega_barrier(P, d.barrier, st);
        if (tracing) trace[3 * i + 2] = sm_clock();
    }
    if (tracing) { trace_cal[1] = sm_clock(); trace_cal[3] = global_timer_ns(); }
    if (P.tp_size > 1 && blockIdx.x == 0 && threadIdx.x == 0) P.sync[96] = st.xchg_base + st.xchg_idx;
}

template <typename T> T* dalloc(size_t n) {
    T* p = nullptr;
    NT_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    NT_CUDA_CHECK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

}  // namespace

#ifndef NT_CUSIM   // everything below talks to the CUDA runtime; the CPU emulation (tests/cusim) stops here

DecodeMega::~DecodeMega() {
    if (abort_host_) cudaFreeHost(abort_host_);
    for (size_t r = 0; r < peer_maps_.size(); r++)
        if (peer_maps_[r] && (int)r != tp_rank_) cudaIpcCloseMemHandle(peer_maps_[r]);
    for (void* p : {(void*)phases_dev_, (void*)hid_[0], (void*)hid_[1], (void*)q_, (void*)k_, (void*)v_, (void*)attn_, (void*)act_,
                    (void*)scratch_, (void*)xq_h_, (void*)xq_a_, (void*)xq_i_, (void*)sync_, xchg_, (void*)cnt_quant_, (void*)cnt_attn_, (void*)cnt_norm_, (void*)ssq_, (void*)trace_})
        if (p) cudaFree(p);
}

bool DecodeMega::build(const MegaModelView& mv) {
    auto fail = [&](const std::string& w) { why_ = w; return false; };
    hidden_ = mv.hidden; nh_ = mv.nh; hd_ = mv.hd; inter_ = mv.inter; tp_rank_ = mv.tp_rank; tp_size_ = mv.tp_size;
    if (hidden_ <= 0 || nh_ <= 0 || hd_ <= 0 || inter_ <= 0 || tp_size_ < 1 || tp_size_ > MEGA_MAX_TP) return fail("bad dimensions");
    int dev = 0, sms = 0, coop = 0;
    NT_CUDA_CHECK(cudaGetDevice(&dev));
    NT_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NT_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return fail("device does not support cooperative launches");
    NT_CUDA_CHECK(cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MEGA_DYN_SMEM));
    int per_sm = 0;
    NT_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel, NTHREADS, MEGA_DYN_SMEM));
    if (per_sm < 1) return fail("decode_step_kernel does not fit one CTA per SM");
    grid_ = sms;

    // ---- buffers ----
    const int qdim = mv.nh * mv.hd, kvdim = mv.nkv * mv.hd;
    hid_[0] = dalloc<float>(hidden_); hid_[1] = dalloc<float>(hidden_);
    q_ = dalloc<float>(qdim); k_ = dalloc<float>(kvdim); v_ = dalloc<float>(kvdim); attn_ = dalloc<float>(qdim);
    act_ = dalloc<float>(inter_);
    xq_h_ = dalloc<int8_t>(xq_bytes(hidden_)); xq_a_ = dalloc<int8_t>(xq_bytes(qdim)); xq_i_ = dalloc<int8_t>(xq_bytes(inter_));
    sync_ = dalloc<unsigned>(MEGA_SYNC_WORDS);
    const size_t slot_floats = (size_t)2 * tp_size_ * hidden_;
    NT_CUDA_CHECK(cudaMalloc(&xchg_, slot_floats * sizeof(float) + (size_t)(tp_size_ + 1) * 32 * sizeof(unsigned)));
    NT_CUDA_CHECK(cudaMemset(xchg_, 0, slot_floats * sizeof(float) + (size_t)(tp_size_ + 1) * 32 * sizeof(unsigned)));

    cnt_quant_ = dalloc<unsigned>((size_t)inter_ / 32 + 1);
    cnt_attn_ = dalloc<unsigned>((size_t)mv.nh + 1);
    cnt_norm_ = dalloc<unsigned>((size_t)hidden_ / 32 + 1);
    ssq_ = dalloc<float>((size_t)hidden_ / 32 + 1);
    if (const char* f = getenv("NT_B200_MEGA_FUSE")) fuse_ = atoi(f);
    MegaBuffers B;
    B.hid[0] = hid_[0]; B.hid[1] = hid_[1]; B.q = q_; B.k = k_; B.v = v_; B.act = act_; B.xq_h = xq_h_; B.xq_a = xq_a_; B.xq_i = xq_i_;
    B.cnt_quant = cnt_quant_; B.cnt_attn = cnt_attn_; B.cnt_norm = cnt_norm_; B.ssq = ssq_;
    if (!mega_make_plan(mv, B, grid_, split_fixed_, fuse_, &plan_, &why_)) return false;
    const std::string bad = mega_check_plan(plan_, grid_, tp_size_);
    if (!bad.empty()) return fail("plan check failed: " + bad);
    scratch_ = dalloc<float>((size_t)mv.nh * plan_.n_splits_max * (mv.hd + 2));
    NT_CUDA_CHECK(cudaMalloc(&phases_dev_, plan_.phases.size() * sizeof(MegaPhase)));
    NT_CUDA_CHECK(cudaMemcpy(phases_dev_, plan_.phases.data(), plan_.phases.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice));

    // ---- kernel parameters ----
    memset(&p_, 0, sizeof(p_));
    p_.phases = phases_dev_; p_.first_gemv = plan_.first_gemv;
    p_.hidden = hidden_; p_.nh = mv.nh; p_.nkv = mv.nkv; p_.hd = mv.hd; p_.gc = plan_.gc; p_.max_seq = mv.max_seq;
    p_.eps = mv.eps; p_.theta = mv.theta; p_.freq_scale = mv.freq_scale; p_.attn_scale = 1.0f / sqrtf((float)mv.hd);
    p_.step = mv.step;
    p_.q = q_; p_.k = k_; p_.v = v_; p_.attn_out = attn_; p_.attn_scratch = scratch_; p_.xq_a = xq_a_;
    p_.n_splits_max = plan_.n_splits_max; p_.split_fixed = plan_.split_fixed; p_.min_split = plan_.min_split; p_.max_split = plan_.max_split;
    p_.attn_smem_off = plan_.attn_smem_off;
    p_.xchg_direct = (plan_.fuse & MEGA_XCHG_DIRECT) ? 1 : 0;
    p_.sync = sync_;
    p_.timeout_ns = 2000000000ull;
    if (const char* t = getenv("NT_B200_MEGA_TIMEOUT_MS")) p_.timeout_ns = (unsigned long long)atoll(t) * 1000000ull;
    p_.tp_rank = tp_rank_; p_.tp_size = tp_size_;
    peer_maps_.assign((size_t)tp_size_, nullptr);
    peer_maps_[(size_t)tp_rank_] = xchg_;
    p_.slots[tp_rank_] = static_cast<float*>(xchg_);
    p_.flags[tp_rank_] = reinterpret_cast<unsigned*>(static_cast<float*>(xchg_) + slot_floats);
    peers_ready_ = (tp_size_ == 1);
    return true;
}

void DecodeMega::export_ipc(void* out64) const {
    static_assert(sizeof(cudaIpcMemHandle_t) == kIpcBytes, "IPC handle size");
    cudaIpcMemHandle_t h;
    NT_CUDA_CHECK(cudaIpcGetMemHandle(&h, xchg_));
    memcpy(out64, &h, kIpcBytes);
}

void DecodeMega::import_peers(const void* handles) {
    const size_t slot_floats = (size_t)2 * tp_size_ * hidden_;
    for (int r = 0; r < tp_size_; r++) {
        if (r == tp_rank_) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * kIpcBytes, kIpcBytes);
        void* p = nullptr;
        NT_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        peer_maps_[(size_t)r] = p;
        p_.slots[r] = static_cast<float*>(p);
        p_.flags[r] = reinterpret_cast<unsigned*>(static_cast<float*>(p) + slot_floats);
    }
    peers_ready_ = true;
}

void DecodeMega::launch(bool with_head, cudaStream_t s) {
    NT_CHECK(phases_dev_ != nullptr, "DecodeMega::launch before build");
    NT_CHECK(peers_ready(), "DecodeMega: tensor-parallel peers not mapped (import_peers)");
    // barrier counter and go word restart at 0 every launch; the abort word and the exchange sequence persist
    NT_CUDA_CHECK(cudaMemsetAsync(sync_, 0, 64 * sizeof(unsigned), s));
    MegaParams p = p_;
    p.n_phases = n_phases(with_head);
    if (const char* mp = getenv("NT_B200_MEGA_MAX_PHASES"))      // bisect aid: stop after k phases, then nt_model_debug_read
        p.n_phases = std::max(1, std::min(p.n_phases, atoi(mp)));
    p.trace = trace_on_ ? trace_ : nullptr;     // laid out for the full program; a body-only launch fills a prefix per CTA
    p.trace_stride = (int)plan_.phases.size() * 3 + 4;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid_); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = MEGA_DYN_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident, or the launch fails loudly instead of deadlocking
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    NT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, decode_step_kernel, p));
    count_launch();
}

void DecodeMega::set_trace(bool on) {
    trace_on_ = on;
    if (on && !trace_) trace_ = dalloc<unsigned long long>((size_t)MEGA_TRACE_CTAS * (plan_.phases.size() * 3 + 4));
}

size_t DecodeMega::read_trace(unsigned long long* out_host, size_t cap) const {
    const size_t n = trace_ ? (size_t)MEGA_TRACE_CTAS * (plan_.phases.size() * 3 + 4) : 0;
    if (out_host && n) NT_CUDA_CHECK(cudaMemcpy(out_host, trace_, sizeof(unsigned long long) * std::min(n, cap), cudaMemcpyDeviceToHost));
    return n;
}

void DecodeMega::enqueue_abort_read(cudaStream_t s) {
    if (!abort_host_) { NT_CUDA_CHECK(cudaMallocHost(&abort_host_, 8 * sizeof(unsigned))); memset(abort_host_, 0, 8 * sizeof(unsigned)); }
    NT_CUDA_CHECK(cudaMemcpyAsync(abort_host_, sync_ + 64, 5 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
}

void DecodeMega::check_abort() {
    if (!abort_host_) return;
    const unsigned* v = abort_host_;
    if (v[0]) {
        fprintf(stderr, "decode_step_kernel (rank %d): a %s wait timed out at barrier #%u of the launch, CTA %u: waiting for %u, saw %u\n", tp_rank_,
                v[0] == 2 ? "tensor-parallel exchange" : "grid-barrier", v[1], v[2], v[3], v[4]);
        NT_CHECK(false, "decode megakernel aborted");
    }
}

const float* DecodeMega::debug_buffer(const char* name, size_t* count) const {
    const std::string n = name ? name : "";
    auto ret = [&](const float* p, size_t c) { if (count) *count = c; return p; };
    if (n == "hid0") return ret(hid_[0], (size_t)hidden_);
    if (n == "hid1") return ret(hid_[1], (size_t)hidden_);
    if (n == "q") return ret(q_, (size_t)nh_ * hd_);
    if (n == "k") return ret(k_, (size_t)(p_.nkv) * hd_);
    if (n == "v") return ret(v_, (size_t)(p_.nkv) * hd_);
    if (n == "attn") return ret(attn_, (size_t)nh_ * hd_);
    if (n == "act") return ret(act_, (size_t)inter_);
    if (n == "slots") return ret(static_cast<const float*>(xchg_), (size_t)2 * tp_size_ * hidden_);
    if (count) *count = 0;
    return nullptr;
}

#endif  // NT_CUSIM

}}  // namespace nt::b200
