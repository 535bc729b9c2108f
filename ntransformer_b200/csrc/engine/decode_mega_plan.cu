// decode_mega_plan.cu — host side of the persistent decode kernel (decode_mega.h): the per-token program (mega_make_plan),
// its structural checker (mega_check_plan) and the CPU replay of every GEMV phase's TMA schedule (mega_check_gemv_schedule).
// Pure host functions without CUDA calls: unit-tested on the CPU (tests/test_mega_plan.py) and reused by the CPU emulation
// harness (tests/cusim/mega_sim.cpp).
#include "decode_mega_sched.h"
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>

namespace nt { namespace b200 {

namespace {
constexpr int AW = MEGA_ATTN_WARPS;
int fmt_of(DType dt) {
    return dt == DType::Q4_K_M ? 0 : dt == DType::Q5_K ? 1 : dt == DType::Q6_K ? 2 : dt == DType::Q8_0 ? 3 : dt == DType::Q4_0 ? 4 : -1;
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
        const int need = std::max((mv.max_seq + split_fixed - 1) / split_fixed, MEGA_COMPAT_MIN_SPLIT);
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

}}  // namespace nt::b200
