// decode_mega.h — the whole decode step (all layers of one token) as ONE persistent kernel.
//
// STATUS: written at the end of round 1 after the round's GPU budget was spent.  It compiles for sm_100a, its host-side
// plan is unit-tested on the CPU and its device code runs — as is, compiled with g++ — on a CPU emulator against the oracle
// (tests/cusim, tests/test_mega_sim.py), but it has NOT run on hardware yet: it is opt-in (NT_B200_MEGAKERNEL=1 or
// nt_model_use_megakernel) and the graph of fused launches in model.cu stays the default.
//
// Why: profiles/r01_launches_70b_summary.txt — the per-layer launches of the decode graph (reference:
// Attention::forward attention.cpp:120-211, FFN::forward ffn.cpp:85-134, called from transformer.cpp:604-669) cost
// 138 us/layer on the 70B Q4_K_M model against 80 us at the HBM roofline.  The gap is fixed cost per launch: every GEMV
// launch drains its TMA rings, ramps up again and pays a grid-wide tail, and the 17-22 us q/k/v and o launches only
// reach 0.3-0.4 of peak; under 8-way tensor parallelism (1/8 of the bytes per launch) the same fixed costs plus 160
// NCCL all-reduces per token cap the step at 119 tok/s.
//
// Design: one cooperative kernel of 148 CTAs x 12 warps interprets a host-built list of phases
//     [norm+quantise] B [q/k/v GEMV] B [RoPE + KV write + split attention] B [combine+quantise] B [o GEMV -> slots] X
//     [norm+quantise] B [gate/up GEMV, SwiGLU] B [quantise] B [down GEMV -> slots] X         (x n_layers)
//     [norm+quantise] B [LM head GEMV]
// B = grid barrier (one atomic arrive per CTA + spin on an L2 word), X = grid barrier + tensor-parallel exchange.
//   * the GEMV phases are the chunk-stationary TMA/dp4a data path of gemv_kquant.cu (same gemv_kq_device.cuh inner
//     loops, same row-group schedule, same combine order => bit-identical outputs), but the per-warp TMA rings are primed
//     for the NEXT GEMV phase before the barrier that precedes it, so HBM keeps streaming weights across barriers and
//     through the small phases;
//   * the residual stream is never accumulated in place: o-proj / down-proj write their partial result into "slots"
//     [parity][source rank][hidden] on EVERY tensor-parallel peer (plain stores to IPC-mapped peer memory over NVLink),
//     the X barrier adds a flag round trip between the ranks' master CTAs, and the following norm phase computes
//     hidden' = hidden + sum_r slot[r] in rank order on every rank (deterministic, identical on all ranks).  With one
//     rank this degenerates to hidden += y.  No NCCL call on the per-layer path;
//   * every spin has a time-out that raises an abort flag instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include "../kernels_internal.h"

namespace nt { namespace b200 {

constexpr int MEGA_WARPS = 12;              // warps per CTA (384 threads, one CTA per SM)
constexpr int MEGA_MAX_STAGES = 4;          // TMA ring depth per warp
constexpr int MEGA_MAX_TP = 8;
constexpr int MEGA_ATTN_WARPS = 8;          // warps that work in the attention phase (attention.cu AW)
constexpr int MEGA_SYNC_WORDS = 4 * 32;     // counter, go, abort, exchange sequence: one 128-byte line each
constexpr int MEGA_COMPAT_MIN_SPLIT = 32;   // compat split rule: keys per context slice at least as in the graph path (attention.cu DYN_MIN_SPLIT)
constexpr int MEGA_TRACE_CTAS = 4;          // CTAs that record the phase timeline when tracing is on

enum MegaPhaseKind : int { MPH_NORM_XQ = 0, MPH_QUANT = 1, MPH_GEMV = 2, MPH_ATTN = 3, MPH_COMBINE = 4, MPH_REDUCE_XQ = 5 };
enum MegaBarrierKind : int { MBAR_NONE = 0, MBAR_GRID = 1, MBAR_EXCHANGE = 2 };
enum MegaEpilogue : int { MEP_STORE = 0, MEP_SWIGLU = 2, MEP_SLOT = 3 };
enum MegaFuse : int { MEGA_FUSE_QUANT = 1, MEGA_FUSE_COMBINE = 2, MEGA_FUSE_NORM = 4 /* single rank only */,
                      MEGA_DEFER_RMS = 8 /* always on under tensor parallelism */,
                      MEGA_OVERLAP_ATTN = 16 /* needs MEGA_FUSE_COMBINE: stream the o-projection's weights during attention */,
                      MEGA_SPLIT_TAIL = 32 /* cut a partly filled last round into 1- or 2-row stages spread over all warp slots */,
                      MEGA_L2_PREFETCH = 64 /* experiment for wide tensor parallelism: pull the NEXT GEMV phase's rows into L2 */,
                      MEGA_XCHG_DIRECT = 128 /* experiment: every CTA signals every rank's arrival counter itself (no master hop) */ };

struct MegaMat {
    const uint8_t* W;
    float* y;
    long long pitch;     // bytes between rows
    int out;             // rows
    int groups;          // ceil(out / 4)
    int fmt;             // 0 Q4_K, 1 Q5_K, 2 Q6_K, 3 Q8_0 (gemv_kq_device.cuh Fmt<>)
    int pad_;
};

// One step of the per-token program.  A flat POD: the kernel copies it into shared memory with 32-bit loads.
struct MegaPhase {
    int kind;            // MegaPhaseKind
    int barrier;         // MegaBarrierKind executed after the phase
    int prime;           // index of the GEMV phase whose TMA rings are primed at the end of this phase, or -1
    int layer;
    // ---- MPH_GEMV ----
    MegaMat mat[3];
    int n_mat, K, NB, NC;
    int n_seg, total_groups, epilogue, warps;
    int gpc, stages, slot_bytes, slot_parity;
    int n_rounds, full_rounds, tail_groups, tail_nr;   // lock-step rounds; the last one may be cut into tail_nr-row stages (MEGA_SPLIT_TAIL)
    const int8_t* xq;           // pre-quantised activations (kernels_internal.h "xq")
    // ---- MPH_NORM_XQ: h = hid_in (+ sum_r slot[pending_parity][r]);  hid_out <- h;  xq_out <- quantise(rmsnorm(h) * norm_w)
    //      MPH_QUANT:   xq_out <- quantise(x[0..n))
    //      MPH_REDUCE_XQ (MEGA_DEFER_RMS): like MPH_NORM_XQ but one warp per 32-element block, no full-vector pass: hid_out <- h,
    //                   ssq_out[block] <- sum(h^2), xq_out <- quantise(h * norm_w); the consuming GEMV applies 1/rms (ssq_in)
    const float* norm_w;
    const float* hid_in;
    float* hid_out;             // null when nothing is pending
    const float* x;
    int8_t* xq_out;
    int pending_parity;         // -1: no slots to add
    int n;
    // ---- MPH_ATTN / MPH_COMBINE: this layer's KV cache ----
    void* kc;
    void* vc;
    // ---- producer-side fusions (MEGA_FUSE_*): "last arriver" counters, zero between uses ----
    //   MPH_GEMV + SwiGLU: the warp that completes the last row-group of a 32-row block quantises the block into xq_out
    //                      (x = the activation vector, n = its length) => no MPH_QUANT phase, one barrier less;
    //   MPH_ATTN:          the unit that completes the last split of a head group merges the splits and emits xq_a
    //                      => no MPH_COMBINE phase, one barrier less.
    //   MPH_GEMV + slots, one rank only (MEGA_FUSE_NORM): the warp that completes a 32-row block adds it to the residual
    //                      stream (hid_in + slot -> hid_out), stores the block's sum of squares in ssq_out and quantises
    //                      h * norm_w (the NEXT norm's weights, without the 1/rms factor) into xq_out; the consuming GEMV
    //                      (ssq_in != null) multiplies its results by rsqrt(sum(ssq_in) / hidden + eps)
    //                      => no MPH_NORM_XQ phase, two barriers less per layer.
    // MEGA_L2_PREFETCH: the matrices of the GEMV phase after this one (whole allocations, 16-byte multiples); every CTA prefetches
    // its 1/grid slice into L2 while this phase computes
    const uint8_t* pf_ptr[3];
    unsigned long long pf_bytes[3];
    unsigned* cnt;
    float* ssq_out;
    const float* ssq_in;
    int fuse;
    int pad_;
};
static_assert(sizeof(MegaPhase) % 4 == 0 && sizeof(MegaPhase) <= 4 * MEGA_WARPS * 32, "MegaPhase is copied by one CTA-wide pass of 32-bit loads");

struct MegaParams {
    const MegaPhase* phases;
    int n_phases, first_gemv;
    int hidden, nh, nkv, hd, gc, max_seq;
    float eps, theta, freq_scale, attn_scale;
    const int* step;            // [0] token, [1] position (device)
    float *q, *k, *v, *attn_out, *attn_scratch;
    int8_t* xq_a;               // attention output in xq form (input of the o-projection)
    int n_splits_max;           // stride of the split dimension in attn_scratch
    int split_fixed;            // > 0: the graph path's rule (ctx cut into this many splits); 0: adaptive
    int min_split, max_split;   // adaptive rule: keys per split; max_split also sizes the score area in shared memory
    int attn_smem_off;          // byte offset of the attention phase's scratch inside the dynamic shared memory (0 = aliases the rings)
    int xchg_direct;            // MEGA_XCHG_DIRECT: the exchange is one remote atomic per CTA and rank instead of a master round trip
    unsigned* sync;             // MEGA_SYNC_WORDS words
    unsigned long long timeout_ns;
    int tp_rank, tp_size;
    // Optional timeline (null = off): CTAs 0..MEGA_TRACE_CTAS-1 store the SM clock at [cta][phase][0 start, 1 work done,
    // 2 barrier passed] followed by 4 calibration values (clock and %globaltimer at kernel start / end).  tools/mega_trace.py turns it into time per phase kind and time spent waiting in barriers.
    unsigned long long* trace;
    int trace_stride;               // values per CTA (3 x phases of the full program + 4)
    int pad2_;
    float* slots[MEGA_MAX_TP];      // rank r's slot buffer [2][tp_size][hidden] as mapped into this process
    unsigned* flags[MEGA_MAX_TP];   // rank r's flag words: one 128-byte line per source rank, then one line with its arrival counter
};

// What the plan builder needs to know about the model (device pointers owned by Model).
struct MegaWeight { const void* ptr; DType dtype; int rows, cols; size_t pitch; };
struct MegaLayerView {
    const float* attn_norm; const float* ffn_norm;
    MegaWeight wq, wk, wv, wo, gate, up, down;
    void* kc; void* vc;
};
struct MegaModelView {
    int hidden = 0, nh = 0, nkv = 0, hd = 0, inter = 0, max_seq = 0, n_layers = 0;
    float eps = 0.f, theta = 0.f, freq_scale = 1.f;
    int tp_rank = 0, tp_size = 1;
    std::vector<MegaLayerView> layers;
    MegaWeight head{};
    const float* out_norm = nullptr;
    float* logits = nullptr;        // where the LM head rows of this rank go
    const int* step = nullptr;
};

// Working buffers the plan refers to (device pointers; fake but aligned addresses in the CPU unit tests).
struct MegaBuffers {
    float* hid[2] = {nullptr, nullptr};
    float *q = nullptr, *k = nullptr, *v = nullptr, *act = nullptr;
    int8_t *xq_h = nullptr, *xq_a = nullptr, *xq_i = nullptr;
    unsigned* cnt_quant = nullptr;      // inter / 32 counters (MEGA_FUSE_QUANT)
    unsigned* cnt_attn = nullptr;       // one counter per attention head group (MEGA_FUSE_COMBINE)
    unsigned* cnt_norm = nullptr;       // hidden / 32 counters (MEGA_FUSE_NORM)
    float* ssq = nullptr;               // hidden / 32 per-block sums of squares of the residual stream (MEGA_FUSE_NORM)
};
struct MegaPlan {
    std::vector<MegaPhase> phases;
    int n_body = 0;                 // phases without the final norm + LM head
    int first_gemv = -1;
    int gc = 0;                     // query heads per attention unit
    int n_splits_max = 0, split_fixed = 0, min_split = 0, max_split = 0;
    int attn_smem_off = 0;          // MEGA_OVERLAP_ATTN: the attention scratch sits above the (smaller) o-projection rings
    int fuse = 0;                   // MegaFuse bits in effect
};
// Pure host functions (no CUDA calls; unit-tested on the CPU through nt_b200_mega_selftest).
bool mega_make_plan(const MegaModelView& mv, const MegaBuffers& B, int grid, int split_fixed, int fuse, MegaPlan* out, std::string* why);
std::string mega_check_gemv_schedule(const MegaPhase& d, int grid, size_t ring_bytes);   // "" = the schedule is consistent
std::string mega_check_plan(const MegaPlan& pl, int grid, int tp_size);                  // "" = structural invariants hold

// Host-side GEMV phase geometry (mirrors launch_fmt/pick_warps of gemv_kquant.cu).  Pure function: unit-tested on the CPU.
struct MegaGemvGeom { int NB, NC, warps, gpc, stages, slot_bytes, mask; bool ok; };
MegaGemvGeom mega_gemv_geom(const int* fmts, int n_mat, int K, size_t ring_bytes);
size_t mega_ring_bytes();       // dynamic shared memory of the kernel

class DecodeMega {
public:
    DecodeMega() = default;
    ~DecodeMega();
    DecodeMega(const DecodeMega&) = delete;
    DecodeMega& operator=(const DecodeMega&) = delete;

    // Builds the phase list and allocates the working buffers.  Returns false (reason in why()) when a shape or dtype is not
    // covered; the caller then keeps the graph path.
    bool build(const MegaModelView& mv);
    const std::string& why() const { return why_; }
    float* embed_out() const { return hid_[0]; }          // the embedding row of the token goes here before launch()
    // Enqueues [reset of the barrier words] + the kernel.  Capturable in a CUDA graph.
    void launch(bool with_head, cudaStream_t s);
    int launches_per_step() const { return 1; }
    // Raises (NT_CHECK) when a spin inside the kernel timed out.  enqueue_abort_read copies the abort words to pinned host memory
    // on the stream (no extra synchronisation); check_abort inspects that copy after the caller has synchronised the stream.
    void enqueue_abort_read(cudaStream_t s);
    void check_abort();

    // Tensor parallel: exchange of the IPC handle of the slot/flag allocation (64 bytes per rank).
    static constexpr int kIpcBytes = 64;
    void export_ipc(void* out64) const;
    void import_peers(const void* handles /* tp_size x 64 bytes, rank order */);
    bool peers_ready() const { return tp_size_ == 1 || peers_ready_; }

    // Debug: plan and buffers.
    const std::vector<MegaPhase>& plan() const { return plan_.phases; }
    int n_phases(bool with_head) const { return with_head ? (int)plan_.phases.size() : plan_.n_body; }
    const float* debug_buffer(const char* name, size_t* count) const;
    void set_split_fixed(int n) { split_fixed_ = n; }      // before build(): use the graph path's split rule (bit-identical attention)
    void set_fuse(int bits) { fuse_ = bits; }              // before build(): MegaFuse bits (default 0: the plain 9-phase program)
    // Phase timeline of the most recent launch ([MEGA_TRACE_CTAS][3 x phases + 4] values); enabling it costs three clock reads per phase.
    void set_trace(bool on);
    size_t read_trace(unsigned long long* out_host, size_t cap) const;     // returns the number of values available

private:
    MegaPlan plan_;
    MegaParams p_{};
    MegaPhase* phases_dev_ = nullptr;
    float* hid_[2] = {nullptr, nullptr};
    float *q_ = nullptr, *k_ = nullptr, *v_ = nullptr, *attn_ = nullptr, *act_ = nullptr, *scratch_ = nullptr;
    int8_t *xq_h_ = nullptr, *xq_a_ = nullptr, *xq_i_ = nullptr;
    unsigned* sync_ = nullptr;
    void* xchg_ = nullptr;             // [slots 2 x tp x hidden floats][flags tp x 32 words], IPC-exportable
    std::vector<void*> peer_maps_;
    int hidden_ = 0, nh_ = 0, hd_ = 0, inter_ = 0, tp_rank_ = 0, tp_size_ = 1, grid_ = 0;
    int split_fixed_ = 0, fuse_ = 0;
    unsigned *cnt_quant_ = nullptr, *cnt_attn_ = nullptr, *cnt_norm_ = nullptr;
    unsigned long long* trace_ = nullptr;
    unsigned* abort_host_ = nullptr;   // pinned copy of the abort words
    bool trace_on_ = false;
    float* ssq_ = nullptr;
    bool peers_ready_ = false;
    std::string why_;
};

}}  // namespace nt::b200
