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
// attention phase.  File layout: memory-model helpers and barriers, the GEMV phase (producer cursor shared with the host-side
// schedule check), norm / quantise / reduce phases, attention, the kernel, then the pure host functions (plan builder and
// checkers, also compiled for the CPU emulator in tests/cusim) and, last, the CUDA-runtime-facing DecodeMega class.
#include "decode_mega.h"
#include "../ring.cuh"
#include "../xquant.cuh"
#include "../gemv_kq_device.cuh"
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

constexpr int NTHREADS = MEGA_WARPS * 32;
constexpr int AW = MEGA_ATTN_WARPS;
constexpr int KU = 4;                                                        // K / V rows in flight per warp in the attention phase
constexpr size_t MEGA_STATIC_SMEM = 4096;                                   // upper bound of the kernel's static __shared__
constexpr size_t MEGA_DYN_SMEM = 227 * 1024 - MEGA_STATIC_SMEM;             // TMA rings; aliased by the attention scratch

// ---- memory-model helpers -----------------------------------------------------------------------------------------
#ifdef NT_CUSIM   // CPU emulation (tests/cusim): C++ atomics; the acquire loads yield so that spinning threads let others run
inline unsigned ld_acquire_gpu(const unsigned* p) { cusim::yield("spin (grid barrier word)"); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned ld_relaxed_sys(const unsigned* p) { cusim::yield("spin (peer flag)"); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned ld_relaxed_gpu(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void red_release_gpu_add(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline void red_relaxed_gpu_add(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline void st_release_gpu(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void st_relaxed_sys(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
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
        if (xchg) {
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

// ---- GEMV phase: producer cursor ----------------------------------------------------------------------------------------
// Identical schedule to gemv_kq_kernel: in round r CTA b handles row-groups (r * grid + b) * gpc + [0, gpc); warp w holds
// chunk (w % NC) of row-group slot (w / NC).  Flattened stage index = round * n_seg + seg.
struct Producer {
    int issued;        // stages issued so far in the phase (uniform across the warp)
    int round, seg;    // round and segment of the next stage to fetch
};

__host__ __device__ __forceinline__ int blk_bytes(int fmt) { return fmt == 1 ? 176 : fmt == 2 ? 210 : fmt == 3 ? 272 : 144; }   // 0 Q4_K, 4 Q4_0: 144

__host__ __device__ __forceinline__ void locate(const MegaPhase& d, int g, int seg, int& mi, int& gl) {
    if (d.n_seg == 2) { mi = seg; gl = g; return; }
    mi = 0;
    while (mi + 1 < d.n_mat && g >= d.mat[mi].groups) { g -= d.mat[mi].groups; mi++; }
    gl = g;
}

// What the stage (round, seg) of warp slot (cta, gsub) holds: nrows rows, starting at row0, of chunk `chunk` of matrix mi.
// The one function both the device producer, the device consumer and the host-side schedule check derive the schedule from.
//   full rounds:  row-group (round * grid + cta) * gpc + gsub, all RG rows;
//   tail round (MEGA_SPLIT_TAIL, tail_nr < RG): the remaining tail_groups row-groups are dealt out tail_nr rows at a time over
//   the slots in slot order, so a partly filled last round costs tail_nr / RG of a round instead of a whole one.
struct StageRef {
    bool empty;            // nothing to do: the stage only completes its mbarrier phase
    int mi, gl, row0, nrows;
    int blkb;              // bytes per 256 weights of the matrix's format
    uint32_t bytes;        // bytes copied per row (multiple of 16)
    long long src_off;     // offset of row row0's part inside the matrix
};
__host__ __device__ __forceinline__ StageRef stage_ref(const MegaPhase& d, int grid, int cta, int gsub, int round, int seg, int chunk,
                                                       int nbc) {
    StageRef r;
    r.empty = true; r.mi = 0; r.gl = 0; r.row0 = 0; r.nrows = RG; r.blkb = 0; r.bytes = 0; r.src_off = 0;
    int g, row_sub = 0, nr = RG;
    if (round < d.full_rounds || d.tail_nr >= RG) {
        g = (round * grid + cta) * d.gpc + gsub;
        if (g >= d.total_groups) return r;
    } else {
        nr = d.tail_nr;
        const int per = RG / nr, u = cta * d.gpc + gsub;
        if (u >= d.tail_groups * per) return r;
        g = d.full_rounds * grid * d.gpc + u / per;
        row_sub = u % per;
    }
    locate(d, g, seg, r.mi, r.gl);
    const MegaMat& m = d.mat[r.mi];
    r.row0 = r.gl * RG + row_sub * nr;
    if (r.row0 >= m.out) return r;                           // a slice of a ragged last group that holds no row
    r.empty = false;
    r.nrows = nr;
    r.blkb = blk_bytes(m.fmt);
    r.bytes = ((uint32_t)(nbc * r.blkb) + 15u) & ~15u;       // a 210-byte tail may spill into row padding (host-checked)
    r.src_off = (long long)chunk * (BS * r.blkb) + (long long)r.row0 * m.pitch;
    return r;
}
__host__ __device__ __forceinline__ void producer_advance(const MegaPhase& d, Producer& pr) {
    if (++pr.seg == d.n_seg) { pr.seg = 0; pr.round++; }
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
        r.split_len = (ctx + P.split_fixed - 1) / P.split_fixed;
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

int fmt_of(DType dt) {
    return dt == DType::Q4_K_M ? 0 : dt == DType::Q5_K ? 1 : dt == DType::Q6_K ? 2 : dt == DType::Q8_0 ? 3 : dt == DType::Q4_0 ? 4 : -1;
}

template <typename T> T* dalloc(size_t n) {
    T* p = nullptr;
    NT_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    NT_CUDA_CHECK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

}  // namespace

size_t mega_ring_bytes() { return MEGA_DYN_SMEM; }

// Mirrors launch_fmt / pick_warps of gemv_kquant.cu with the CTA width fixed at MEGA_WARPS: the widest multiple of NC that
// affords two ring stages inside `ring_bytes`, ring depth up to MEGA_MAX_STAGES.
MegaGemvGeom mega_gemv_geom(const int* fmts, int n_mat, int K, size_t ring_bytes) {
    MegaGemvGeom g{};
    g.ok = false;
    if (n_mat < 1 || n_mat > 3 || K <= 0 || K % 256 != 0) return g;
    g.NB = K / 256;
    g.NC = (g.NB + BS - 1) / BS;
    if (g.NC > MEGA_WARPS) return g;
    g.mask = 0;
    for (int i = 0; i < n_mat; i++) {
        if (fmts[i] < 0 || fmts[i] > 4) return g;
        g.mask |= 1 << fmts[i];
    }
    g.slot_bytes = RG * BS * max_blk(g.mask);
    int best = 0;
    for (int min_stages = 2; min_stages >= 1 && !best; min_stages--)
        for (int w = g.NC; w <= MEGA_WARPS; w += g.NC)
            if ((size_t)w * min_stages * g.slot_bytes <= ring_bytes) best = w;
    if (!best) return g;
    g.warps = best;
    g.gpc = best / g.NC;
    g.stages = (int)std::min<size_t>(MEGA_MAX_STAGES, ring_bytes / ((size_t)best * g.slot_bytes));
    g.ok = g.stages >= 1;
    return g;
}

// Pure host function (no CUDA calls): the per-token program for `mv` with working buffers `B` on a grid of `grid` CTAs.
bool mega_make_plan(const MegaModelView& mv, const MegaBuffers& B, int grid, int split_fixed, int fuse, MegaPlan* out, std::string* why) {
    auto fail = [&](const std::string& w) { if (why) *why = w; return false; };
    MegaPlan& pl = *out;
    pl = MegaPlan{};
    const int hidden = mv.hidden, inter = mv.inter;
    if (mv.n_layers < 1 || (int)mv.layers.size() != mv.n_layers) return fail("no layers");
    if (mv.tp_size < 1 || mv.tp_size > MEGA_MAX_TP || mv.tp_rank < 0 || mv.tp_rank >= mv.tp_size) return fail("tp_size must be 1..8");
    if (mv.nkv < 1 || mv.nh % mv.nkv != 0) return fail("n_heads % n_kv_heads != 0");
    const int ratio = mv.nh / mv.nkv;
    const int gc = ratio % 8 == 0 ? 8 : ratio % 4 == 0 ? 4 : ratio % 2 == 0 ? 2 : 1;      // attention.cu pick_gc
    if (!((mv.hd == 128 && (gc == 8 || gc == 4)) || (mv.hd == 64 && gc == 4)))
        return fail("attention shape not instantiated (head_dim 128 with 4/8 query heads per KV head, or 64 with 4)");
    if (hidden % 256 != 0 || (mv.nh * mv.hd) % 256 != 0 || inter % 256 != 0) return fail("dimensions must be multiples of 256");
    if (grid < 1) return fail("empty grid");
    pl.gc = gc;
    if ((fuse & MEGA_FUSE_QUANT) && !B.cnt_quant) fuse &= ~MEGA_FUSE_QUANT;
    if ((fuse & MEGA_FUSE_COMBINE) && !B.cnt_attn) fuse &= ~MEGA_FUSE_COMBINE;
    if ((fuse & MEGA_FUSE_NORM) && (!B.cnt_norm || !B.ssq || mv.tp_size != 1 || hidden % 32 != 0)) fuse &= ~MEGA_FUSE_NORM;
    // Under tensor parallelism the norm phase would read tp_size + 1 full vectors per participating CTA: always use the
    // distributed reduce phase there.  At one rank it is optional (it gives up bit-comparability with the graph path).
    if (mv.tp_size > 1 && B.ssq) fuse |= MEGA_DEFER_RMS;
    if ((fuse & MEGA_DEFER_RMS) && !B.ssq) fuse &= ~MEGA_DEFER_RMS;
    pl.fuse = fuse;
    if (!(fuse & MEGA_DEFER_RMS) && hidden / 256 > grid) return fail("hidden / 256 exceeds the grid");   // MPH_NORM_XQ needs them
    const int qdim = mv.nh * mv.hd;

    // ---- attention split geometry ----
    const int n_groups = mv.nh / gc;
    pl.min_split = 64;
    // score area + cross-warp reduction + rotated queries + the token's own k/v must fit the ring area
    const size_t fixed = ((size_t)AW * gc * mv.hd + (size_t)gc * mv.hd) * sizeof(float) + 2 * (size_t)mv.hd * sizeof(__half);
    int max_split = (int)std::min<size_t>(2048, (MEGA_DYN_SMEM - fixed) / ((size_t)gc * sizeof(float)));
    if (split_fixed > 0) {
        const int need = (mv.max_seq + split_fixed - 1) / split_fixed;
        if (need > max_split) return fail("split_fixed: context slice does not fit shared memory");
        pl.n_splits_max = split_fixed;
    } else {
        max_split = std::min(max_split, 256);
        const int cap = std::max(1, grid / n_groups);
        pl.n_splits_max = std::max(cap, (mv.max_seq + max_split - 1) / max_split);
    }
    pl.max_split = max_split;
    pl.split_fixed = split_fixed;
    // MEGA_OVERLAP_ATTN: keep the attention scratch out of the way of the o-projection's (smaller) rings, so that those can
    // be primed at the end of the q/k/v phase and HBM keeps streaming weights through the barrier and the attention phase.
    const size_t attn_bytes = ((fixed + (size_t)gc * max_split * sizeof(float)) + 127) & ~(size_t)127;
    size_t o_ring = MEGA_DYN_SMEM;
    if ((fuse & MEGA_OVERLAP_ATTN) && (fuse & MEGA_FUSE_COMBINE) && attn_bytes + 4 * RG * BS * 144 <= MEGA_DYN_SMEM) {
        pl.attn_smem_off = (int)(MEGA_DYN_SMEM - attn_bytes);
        o_ring = (size_t)pl.attn_smem_off;
    } else {
        fuse &= ~MEGA_OVERLAP_ATTN;
    }
    pl.fuse = fuse;

    // ---- phase list ----
    auto base_phase = [&](int kind, int layer) {
        MegaPhase ph;
        memset(&ph, 0, sizeof(ph));
        ph.kind = kind; ph.barrier = MBAR_GRID; ph.prime = -1; ph.layer = layer; ph.pending_parity = -1;
        return ph;
    };
    auto gemv_phase = [&](int layer, const MegaWeight* const* ws, float* const* ys, int n, int epilogue, const int8_t* xq,
                          int slot_parity, MegaPhase* outp, size_t ring_budget = MEGA_DYN_SMEM) -> bool {
        MegaPhase ph = base_phase(MPH_GEMV, layer);
        GemvMat mats[3];
        int fmts[3];
        const int K = ws[0]->cols;
        for (int i = 0; i < n; i++) {
            if (ws[i]->cols != K) return false;
            mats[i].W = ws[i]->ptr; mats[i].y = ys[i]; mats[i].out = ws[i]->rows; mats[i].dtype = ws[i]->dtype; mats[i].row_pitch = ws[i]->pitch;
            fmts[i] = fmt_of(ws[i]->dtype);
        }
        // same alignment rules as gemv_kq_supported (gemv_kquant.cu), with Q4_0 admitted unconditionally on this opt-in path
        if (K <= 0 || K % 256 != 0 || (K / 256 + BS - 1) / BS > MEGA_WARPS) return false;
        for (int i = 0; i < n; i++) {
            if (fmts[i] < 0 || mats[i].out <= 0) return false;
            const size_t pitch = mats[i].row_pitch ? mats[i].row_pitch : dtype_row_size(mats[i].dtype, (size_t)K);
            if ((reinterpret_cast<uintptr_t>(mats[i].W) & 15) || (pitch & 15)) return false;
            const size_t blk = dtype_row_size(mats[i].dtype, 256);           // bytes per 256 weights
            const int NBq = K / 256, NCq = (NBq + BS - 1) / BS, last = NBq - (NCq - 1) * BS;
            if ((size_t)(NCq - 1) * BS * blk + (((size_t)last * blk + 15) & ~(size_t)15) > pitch) return false;   // 16-byte copy tail
        }
        if (epilogue == MEP_SWIGLU && (n != 2 || ws[0]->rows != ws[1]->rows)) return false;
        const MegaGemvGeom g = mega_gemv_geom(fmts, n, K, ring_budget);
        if (!g.ok) return false;
        int total = 0;
        for (int i = 0; i < n; i++) {
            MegaMat& m = ph.mat[i];
            m.W = static_cast<const uint8_t*>(ws[i]->ptr); m.y = ys[i]; m.out = ws[i]->rows; m.groups = (ws[i]->rows + RG - 1) / RG;
            m.fmt = fmts[i];
            m.pitch = (long long)(ws[i]->pitch ? ws[i]->pitch : dtype_row_size(ws[i]->dtype, (size_t)K));
            total += m.groups;
        }
        ph.n_mat = n; ph.K = K; ph.NB = g.NB; ph.NC = g.NC;
        ph.epilogue = epilogue;
        if (epilogue == MEP_SWIGLU) { ph.n_seg = 2; ph.total_groups = ph.mat[0].groups; }
        else { ph.n_seg = 1; ph.total_groups = total; }
        ph.warps = g.warps; ph.gpc = g.gpc; ph.stages = g.stages; ph.slot_bytes = g.slot_bytes;
        ph.slot_parity = slot_parity;
        // lock-step rounds over grid * gpc warp slots; a partly filled last round is cut into 1- or 2-row stages
        const int slots = grid * g.gpc;
        ph.full_rounds = ph.total_groups / slots;
        ph.tail_groups = ph.total_groups % slots;
        ph.n_rounds = ph.full_rounds + (ph.tail_groups ? 1 : 0);
        ph.tail_nr = RG;
        if ((fuse & MEGA_SPLIT_TAIL) && ph.tail_groups > 0)
            ph.tail_nr = (ph.tail_groups * 4 <= slots) ? 1 : (ph.tail_groups * 2 <= slots) ? 2 : RG;
        ph.xq = xq;
        *outp = ph;
        return true;
    };

    int cur = 0;                 // B.hid[cur] holds the residual stream (before pending slots are added)
    int pending = -1;            // parity of the slots still to be added, -1 none
    bool deferred = false;       // the current xq_h lacks its 1/rms factor: its consumers apply it themselves (ssq_in)
    const bool defer = (fuse & MEGA_DEFER_RMS) != 0;
    auto norm_phase = [&](int layer, const float* w) {
        MegaPhase ph = base_phase(defer ? MPH_REDUCE_XQ : MPH_NORM_XQ, layer);
        ph.norm_w = w; ph.hid_in = B.hid[cur]; ph.xq_out = B.xq_h; ph.pending_parity = pending;
        if (pending >= 0) { ph.hid_out = B.hid[cur ^ 1]; cur ^= 1; pending = -1; }
        if (defer) ph.ssq_out = B.ssq;
        deferred = defer;
        return ph;
    };
    std::vector<MegaPhase>& plan = pl.phases;
    const bool fnorm = (fuse & MEGA_FUSE_NORM) != 0;
    // fold "residual add + next norm" into a slot-epilogue GEMV phase (MEGA_FUSE_NORM, single rank)
    auto fold_norm = [&](MegaPhase& ph, const float* next_norm_w) {
        ph.fuse |= MEGA_FUSE_NORM;
        ph.norm_w = next_norm_w; ph.hid_in = B.hid[cur]; ph.hid_out = B.hid[cur ^ 1]; cur ^= 1;
        ph.xq_out = B.xq_h; ph.cnt = B.cnt_norm; ph.ssq_out = B.ssq;
        deferred = true;
    };
    for (int l = 0; l < mv.n_layers; l++) {
        const MegaLayerView& L = mv.layers[(size_t)l];
        if (!L.attn_norm || !L.ffn_norm) return fail("missing norm weights");
        const float* next_attn_norm = (l + 1 < mv.n_layers) ? mv.layers[(size_t)l + 1].attn_norm : mv.out_norm;
        if (fnorm && !next_attn_norm) return fail("missing norm weights");
        if (!fnorm || l == 0) plan.push_back(norm_phase(l, L.attn_norm));
        MegaPhase ph;
        { const MegaWeight* ws[3] = {&L.wq, &L.wk, &L.wv}; float* ys[3] = {B.q, B.k, B.v};
          if (L.wq.rows != qdim || L.wk.rows != mv.nkv * mv.hd || L.wv.rows != mv.nkv * mv.hd || L.wq.cols != hidden) return fail("attn_q/k/v shape");
          if (!gemv_phase(l, ws, ys, 3, MEP_STORE, B.xq_h, 0, &ph)) return fail("q/k/v weights not on the K-quant TMA path");
          if (deferred) ph.ssq_in = B.ssq;
          plan.push_back(ph); }
        ph = base_phase(MPH_ATTN, l); ph.kc = L.kc; ph.vc = L.vc;
        if (fuse & MEGA_FUSE_COMBINE) { ph.fuse = MEGA_FUSE_COMBINE; ph.cnt = B.cnt_attn; }
        plan.push_back(ph);
        if (!(fuse & MEGA_FUSE_COMBINE)) { ph = base_phase(MPH_COMBINE, l); plan.push_back(ph); }
        { const MegaWeight* ws[1] = {&L.wo}; float* ys[1] = {nullptr};
          if (L.wo.rows != hidden || L.wo.cols != qdim) return fail("attn_output shape");
          if (!gemv_phase(l, ws, ys, 1, MEP_SLOT, B.xq_a, 0, &ph, o_ring)) return fail("attn_output weight not on the K-quant TMA path");
          ph.barrier = MBAR_EXCHANGE;
          if (fnorm) fold_norm(ph, L.ffn_norm); else pending = 0;
          plan.push_back(ph); }
        if (!fnorm) plan.push_back(norm_phase(l, L.ffn_norm));
        { const MegaWeight* ws[2] = {&L.gate, &L.up}; float* ys[2] = {B.act, nullptr};
          if (L.gate.rows != inter || L.up.rows != inter || L.gate.cols != hidden) return fail("ffn_gate/up shape");
          if (!gemv_phase(l, ws, ys, 2, MEP_SWIGLU, B.xq_h, 0, &ph)) return fail("ffn_gate/up weights not on the K-quant TMA path");
          if (fuse & MEGA_FUSE_QUANT) { ph.fuse = MEGA_FUSE_QUANT; ph.cnt = B.cnt_quant; ph.x = B.act; ph.n = inter; ph.xq_out = B.xq_i; }
          if (deferred) ph.ssq_in = B.ssq;
          plan.push_back(ph); }
        if (!(fuse & MEGA_FUSE_QUANT)) { ph = base_phase(MPH_QUANT, l); ph.x = B.act; ph.n = inter; ph.xq_out = B.xq_i; plan.push_back(ph); }
        { const MegaWeight* ws[1] = {&L.down}; float* ys[1] = {nullptr};
          if (L.down.rows != hidden || L.down.cols != inter) return fail("ffn_down shape");
          if (!gemv_phase(l, ws, ys, 1, MEP_SLOT, B.xq_i, 1, &ph)) return fail("ffn_down weight not on the K-quant TMA path");
          ph.barrier = MBAR_EXCHANGE;
          if (fnorm) fold_norm(ph, next_attn_norm); else pending = 1;
          plan.push_back(ph); }
    }
    pl.n_body = (int)plan.size();
    if (!mv.out_norm || !mv.logits) return fail("missing output norm / logits buffer");
    if (!fnorm) plan.push_back(norm_phase(mv.n_layers, mv.out_norm));
    if (mv.head.rows > 0) {
        MegaPhase ph;
        const MegaWeight* ws[1] = {&mv.head}; float* ys[1] = {mv.logits};
        if (mv.head.cols != hidden) return fail("output.weight shape");
        if (!gemv_phase(mv.n_layers, ws, ys, 1, MEP_STORE, B.xq_h, 0, &ph)) return fail("output.weight not on the K-quant TMA path");
        if (deferred) ph.ssq_in = B.ssq;
        ph.barrier = MBAR_NONE;
        plan.push_back(ph);
    } else {
        plan.back().barrier = MBAR_NONE;
    }
    // ring priming: at the end of a phase, start the next GEMV phase's weight stream unless an attention phase (which
    // aliases the ring area) still lies in between; the attention phase itself primes the GEMV that follows it.
    if (fuse & MEGA_L2_PREFETCH) {
        int prev = -1;
        for (int i = 0; i < (int)plan.size(); i++) {
            if (plan[i].kind != MPH_GEMV) continue;
            if (prev >= 0) {
                plan[prev].fuse |= MEGA_L2_PREFETCH;
                for (int m = 0; m < plan[i].n_mat; m++) {
                    plan[prev].pf_ptr[m] = plan[i].mat[m].W;
                    plan[prev].pf_bytes[m] = ((unsigned long long)plan[i].mat[m].out * (unsigned long long)plan[i].mat[m].pitch) & ~15ull;
                }
            }
            prev = i;
        }
    }
    pl.first_gemv = -1;
    const bool overlap = (pl.fuse & MEGA_OVERLAP_ATTN) != 0;
    for (int i = 0; i < (int)plan.size(); i++) {
        if (plan[i].kind == MPH_GEMV && pl.first_gemv < 0) pl.first_gemv = i;
        if (plan[i].kind != MPH_GEMV && plan[i].kind != MPH_ATTN) continue;
        if (plan[i].kind == MPH_ATTN && overlap) continue;       // its successor was primed by the q/k/v phase already
        for (int j = i + 1; j < (int)plan.size(); j++) {
            if (plan[j].kind == MPH_ATTN && !overlap) break;
            if (plan[j].kind == MPH_GEMV) { plan[i].prime = j; break; }
        }
    }
    return true;
}

// Host-side replay of one GEMV phase's schedule on `grid` CTAs with the very cursor functions the kernel uses: every warp's
// producer sequence must equal its consumer sequence, every (matrix, row-group, chunk) must be fetched exactly once, every
// copy must stay inside its row and the ring inside `ring_bytes`.  Returns an empty string when all of that holds.
std::string mega_check_gemv_schedule(const MegaPhase& d, int grid, size_t ring_bytes) {
    char msg[256];
    if (d.kind != MPH_GEMV) return "not a GEMV phase";
    if (d.warps < 1 || d.warps > MEGA_WARPS || d.warps % d.NC != 0 || d.gpc != d.warps / d.NC) return "warps / NC / gpc inconsistent";
    if (d.stages < 1 || d.stages > MEGA_MAX_STAGES) return "ring depth out of range";
    if ((size_t)d.warps * d.stages * d.slot_bytes > ring_bytes) return "rings exceed the dynamic shared memory";
    if (d.NB != d.K / 256 || d.NC != (d.NB + BS - 1) / BS) return "NB / NC inconsistent with K";
    const int slots = grid * d.gpc;
    if (d.full_rounds != d.total_groups / slots || d.tail_groups != d.total_groups % slots ||
        d.n_rounds != d.full_rounds + (d.tail_groups ? 1 : 0)) return "round bookkeeping inconsistent with the grid";
    if (d.tail_nr != 1 && d.tail_nr != 2 && d.tail_nr != RG) return "tail stage height must be 1, 2 or RG rows";
    if (d.tail_nr < RG && d.tail_groups * (RG / d.tail_nr) > slots) return "tail round does not fit the warp slots";
    const int n_total = d.n_rounds * d.n_seg;
    std::vector<std::vector<unsigned char>> seen((size_t)d.n_mat);          // per matrix: [row][chunk]
    for (int i = 0; i < d.n_mat; i++) seen[(size_t)i].assign((size_t)d.mat[i].out * d.NC, 0);
    for (int b = 0; b < grid; b++) {
        for (int w = 0; w < d.warps; w++) {
            const int chunk = w % d.NC, gsub = w / d.NC, nbc = std::min(BS, d.NB - chunk * BS);
            Producer pr;
            pr.issued = 0; pr.round = 0; pr.seg = 0;
            int s = 0;
            for (int round = 0; round < d.n_rounds; round++) {
                const StageRef first = stage_ref(d, grid, b, gsub, round, 0, chunk, nbc);   // what the consumer derives for the round
                for (int seg = 0; seg < d.n_seg; seg++, s++) {
                    if (pr.round != round || pr.seg != seg) {
                        snprintf(msg, sizeof(msg), "cta %d warp %d stage %d: producer (%d,%d) != consumer (%d,%d)", b, w, s, pr.round, pr.seg, round, seg);
                        return msg;
                    }
                    const StageRef sr = stage_ref(d, grid, b, gsub, pr.round, pr.seg, chunk, nbc);
                    if (sr.empty != first.empty || (!sr.empty && (sr.row0 != first.row0 || sr.nrows != first.nrows)))
                        return "segments of one round disagree on their rows";
                    if (!sr.empty) {
                        const MegaMat& m = d.mat[sr.mi];
                        if (d.n_seg == 2 && sr.mi != seg) return "SwiGLU segment does not select its matrix";
                        if (sr.gl < 0 || sr.gl >= m.groups || sr.row0 < 0 || sr.row0 >= m.out) return "rows outside their matrix";
                        if (sr.nrows != RG && sr.nrows != d.tail_nr) return "unexpected stage height";
                        if ((sr.row0 >> 5) != ((sr.row0 + sr.nrows - 1) >> 5)) return "a stage straddles two 32-row blocks";
                        if ((size_t)sr.nrows * BS * sr.blkb > (size_t)d.slot_bytes) return "stage larger than its ring slot";
                        if ((long long)chunk * (BS * sr.blkb) + (long long)sr.bytes > m.pitch) return "copy runs past the row pitch";
                        if ((sr.src_off & 15) || (m.pitch & 15) || (sr.bytes & 15)) return "copy is not 16-byte aligned";
                        for (int r = 0; r < sr.nrows && sr.row0 + r < m.out; r++) {
                            unsigned char& c = seen[(size_t)sr.mi][(size_t)(sr.row0 + r) * d.NC + chunk];
                            if (c) { snprintf(msg, sizeof(msg), "matrix %d row %d chunk %d fetched twice", sr.mi, sr.row0 + r, chunk); return msg; }
                            c = 1;
                        }
                    }
                    producer_advance(d, pr);
                }
            }
            if (s != n_total) return "stage count mismatch";
        }
    }
    for (int i = 0; i < d.n_mat; i++)
        for (size_t j = 0; j < seen[(size_t)i].size(); j++)
            if (!seen[(size_t)i][j]) { snprintf(msg, sizeof(msg), "matrix %d row %zu chunk %zu never fetched", i, j / d.NC, j % d.NC); return msg; }
    return "";
}

// Structural invariants of a whole plan (what the kernel silently relies on).  Returns an empty string when they hold.
std::string mega_check_plan(const MegaPlan& pl, int grid, int tp_size) {
    char msg[256];
    const std::vector<MegaPhase>& ph = pl.phases;
    const int n = (int)ph.size();
    if (n == 0 || pl.n_body <= 0 || pl.n_body > n) return "empty plan";
    if (pl.first_gemv > 1) return "first GEMV phase beyond the initially loaded descriptors (0, 1)";
    int primed = pl.first_gemv;                  // GEMV phase whose rings are currently primed, -1 none
    const float* stream = nullptr;               // buffer that holds the residual stream
    int pending = -1;                            // parity of the slots waiting to be added
    bool deferred = false;                       // the latest norm output lacks its 1/rms factor (fused norm)
    const int8_t* xq_norm = nullptr;             // where the latest norm wrote its quantised output
    for (int i = 0; i < n; i++) {
        const MegaPhase& d = ph[(size_t)i];
        if (i + 1 < n && d.barrier == MBAR_NONE) { snprintf(msg, sizeof(msg), "phase %d: no barrier before phase %d", i, i + 1); return msg; }
        if (d.kind == MPH_GEMV) {
            if (primed != i) { snprintf(msg, sizeof(msg), "GEMV phase %d starts with rings primed for %d", i, primed); return msg; }
            primed = -1;
            // the 1/rms factor of a norm is applied exactly once: by the norm phase itself, or by the consumers of a fused norm
            if ((d.ssq_in != nullptr) != (d.xq == xq_norm && deferred)) { snprintf(msg, sizeof(msg), "GEMV phase %d: 1/rms factor applied twice or never", i); return msg; }
            const std::string e = mega_check_gemv_schedule(d, grid, MEGA_DYN_SMEM);
            if (!e.empty()) { snprintf(msg, sizeof(msg), "GEMV phase %d: %s", i, e.c_str()); return msg; }
            if (d.epilogue == MEP_SLOT) {
                if (d.barrier != MBAR_EXCHANGE) return "slot epilogue without an exchange barrier";
                if (pending >= 0) return "two exchanges without a norm phase in between";
                if (d.fuse & MEGA_FUSE_NORM) {                 // the epilogue adds the slots to the residual stream itself
                    if (tp_size != 1) return "fused norm under tensor parallelism";
                    if (stream && d.hid_in != stream) return "fused norm does not read the current residual stream";
                    if (!d.hid_out || d.hid_out == d.hid_in || !d.ssq_out || !d.cnt || !d.norm_w || !d.xq_out) return "fused norm fields";
                    stream = d.hid_out;
                    xq_norm = d.xq_out;
                    deferred = true;
                } else {
                    pending = d.slot_parity;
                }
            } else if (d.barrier == MBAR_EXCHANGE) return "exchange barrier after a non-slot phase";

        } else if (d.kind == MPH_ATTN) {
            if (primed >= 0) {                                   // allowed only when the scratch sits above the primed rings
                const MegaPhase& g = ph[(size_t)primed];
                if (pl.attn_smem_off <= 0 || (size_t)g.warps * g.stages * g.slot_bytes > (size_t)pl.attn_smem_off) {
                    snprintf(msg, sizeof(msg), "attention phase %d would overwrite rings primed for %d", i, primed);
                    return msg;
                }
            }
        } else if (d.kind == MPH_NORM_XQ || d.kind == MPH_REDUCE_XQ) {
            if (stream && d.hid_in != stream) { snprintf(msg, sizeof(msg), "norm phase %d does not read the current residual stream", i); return msg; }
            if (d.pending_parity != pending) { snprintf(msg, sizeof(msg), "norm phase %d: pending parity %d, expected %d", i, d.pending_parity, pending); return msg; }
            if ((pending >= 0) != (d.hid_out != nullptr)) return "hid_out must be set exactly when slots are pending";
            if (d.hid_out == d.hid_in) return "residual stream updated in place";
            stream = d.hid_out ? d.hid_out : d.hid_in;
            pending = -1;
            xq_norm = d.xq_out;
            deferred = (d.kind == MPH_REDUCE_XQ);
            if (deferred && !d.ssq_out) return "reduce phase without a sum-of-squares buffer";
        }
        if (d.prime >= 0) {
            if (d.prime <= i || d.prime >= n || ph[(size_t)d.prime].kind != MPH_GEMV) return "prime target is not a later GEMV phase";
            if (d.prime > i + 2) return "prime target beyond the descriptor prefetch window (i + 2)";
            if (primed >= 0) return "rings primed twice";
            for (int j = i + 1; j < d.prime; j++)
                if (ph[(size_t)j].kind == MPH_GEMV || (ph[(size_t)j].kind == MPH_ATTN && pl.attn_smem_off <= 0))
                    return "phases between a prime and its GEMV touch the rings";
            primed = d.prime;
        }
    }
    if (primed >= 0) return "plan ends with primed rings";
    (void)tp_size;
    return "";
}

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
    NT_CUDA_CHECK(cudaMalloc(&xchg_, slot_floats * sizeof(float) + (size_t)tp_size_ * 32 * sizeof(unsigned)));
    NT_CUDA_CHECK(cudaMemset(xchg_, 0, slot_floats * sizeof(float) + (size_t)tp_size_ * 32 * sizeof(unsigned)));

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
