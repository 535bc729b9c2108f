// mega_sim.cpp — runs the persistent decode kernel's OWN source on the CPU emulator (cusim.h).  TEST INFRASTRUCTURE.
//
// Build (tests/cusim/Makefile): g++ -DNT_CUSIM, this file #includes decode_megakernel.cu, links libnt_b200.so for the host
// helpers it calls (GGUF reader, gemv_kq_supported).  A simulated model is tp_size "GPUs" in one process: each rank owns its
// weight shards, KV cache, working buffers and plan (mega_make_plan, the production builder), the ranks' slot/flag buffers
// are cross-linked like IPC-mapped peer memory, and one step launches tp_size emulated grids concurrently.
#include "../../ntransformer_b200/csrc/engine/decode_mega.h"     // before cusim.h: launch_k names cudaLaunchConfig_t::gridDim
#include "../../ntransformer_b200/csrc/engine/gguf.h"
#include <algorithm>
#include <memory>
#include <string>
#include "cusim.h"
#include "../../ntransformer_b200/csrc/engine/decode_megakernel.cu"

using namespace nt::b200;

namespace {

size_t round16(size_t v) { return (v + 15) & ~(size_t)15; }

struct HostBuf {
    std::vector<uint8_t> raw;
    void* ptr = nullptr;
    void alloc(size_t bytes) {
        raw.assign(bytes + 256, 0);
        ptr = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(raw.data()) + 255) & ~(uintptr_t)255);
    }
    template <typename T> T* as() { return static_cast<T*>(ptr); }
};

struct Rank {
    std::vector<std::unique_ptr<HostBuf>> owned;
    MegaModelView mv;
    MegaBuffers B;
    MegaPlan plan;
    MegaParams P{};
    HostBuf hid[2], q, k, v, attn, act, xq_h, xq_a, xq_i, scratch, sync, xchg, logits, step, cnt_quant, cnt_attn, cnt_norm, ssq;
    std::vector<HostBuf> kc, vc;

    void* keep(size_t bytes) {
        owned.emplace_back(new HostBuf());
        owned.back()->alloc(bytes);
        return owned.back()->ptr;
    }
};

struct Sim {
    GGUFFile file;
    ModelConfig cfg;
    int tp = 1, grid = 4, copy_delay = 0, vocab_l = 0;
    std::vector<std::unique_ptr<Rank>> ranks;
};

// this rank's shard of tensor `name` in host memory (Model::upload's rules: rows / columns at block boundaries, column shards
// re-pitched to 16 bytes)
MegaWeight shard(Sim& S, Rank& R, const std::string& name, int split, int rank) {
    const GGUFTensorInfo* ti = S.file.find(name);
    NT_CHECK(ti != nullptr, ("tensor not found: " + name).c_str());
    const uint8_t* src = static_cast<const uint8_t*>(S.file.data(*ti));
    const nt::DType dt = ti->dtype;
    const int cols = (int)ti->shape[0], rows = ti->shape.size() > 1 ? (int)ti->shape[1] : 1;
    const size_t row_bytes = nt::dtype_row_size(dt, (size_t)cols);
    MegaWeight w{};
    w.dtype = dt;
    if (split == 0 || S.tp == 1) {
        void* d = R.keep(ti->nbytes);
        memcpy(d, src, ti->nbytes);
        w.ptr = d; w.rows = rows; w.cols = cols; w.pitch = row_bytes;
    } else if (split == 1) {
        const int per = (rows + S.tp - 1) / S.tp, r0 = std::min(rows, rank * per), r1 = std::min(rows, r0 + per);
        void* d = R.keep(std::max<size_t>((size_t)(r1 - r0) * row_bytes, 16));
        memcpy(d, src + (size_t)r0 * row_bytes, (size_t)(r1 - r0) * row_bytes);
        w.ptr = d; w.rows = r1 - r0; w.cols = cols; w.pitch = row_bytes;
    } else {
        const int c_l = cols / S.tp;
        const size_t shard_row = nt::dtype_row_size(dt, (size_t)c_l), pitch = round16(shard_row);
        uint8_t* d = static_cast<uint8_t*>(R.keep(pitch * rows));
        for (int r = 0; r < rows; r++) memcpy(d + (size_t)r * pitch, src + (size_t)r * row_bytes + (size_t)rank * shard_row, shard_row);
        w.ptr = d; w.rows = rows; w.cols = c_l; w.pitch = pitch;
    }
    return w;
}

}  // namespace

extern "C" {

void* mega_sim_create(const char* gguf_path, int max_seq, int tp_size, int grid, int split_fixed, int fuse, int copy_delay, char* msg,
                      size_t cap) {
    auto say = [&](const std::string& m) { if (msg && cap) snprintf(msg, cap, "%s", m.c_str()); };
    auto S = std::make_unique<Sim>();
    if (!S->file.open(gguf_path)) { say("cannot open gguf"); return nullptr; }
    S->cfg = S->file.config();
    if (S->cfg.max_seq_len > max_seq) S->cfg.max_seq_len = max_seq;
    const ModelConfig& c = S->cfg;
    S->tp = tp_size; S->grid = grid; S->copy_delay = copy_delay;
    if (c.n_heads % tp_size || c.n_kv_heads % tp_size || c.intermediate_size % tp_size) { say("shape does not divide by tp"); return nullptr; }
    const int nh = c.n_heads / tp_size, nkv = c.n_kv_heads / tp_size, inter = c.intermediate_size / tp_size, hd = c.head_dim;
    S->vocab_l = (c.vocab_size + tp_size - 1) / tp_size;
    for (int r = 0; r < tp_size; r++) {
        S->ranks.emplace_back(new Rank());
        Rank& R = *S->ranks.back();
        MegaModelView& mv = R.mv;
        mv.hidden = c.hidden_size; mv.nh = nh; mv.nkv = nkv; mv.hd = hd; mv.inter = inter; mv.max_seq = c.max_seq_len;
        mv.n_layers = c.n_layers; mv.eps = c.norm_eps; mv.theta = c.rope_theta; mv.freq_scale = c.rope_freq_scale;
        mv.tp_rank = r; mv.tp_size = tp_size;
        const size_t kv_bytes = (size_t)c.max_seq_len * nkv * hd * 2;
        R.kc.resize((size_t)c.n_layers); R.vc.resize((size_t)c.n_layers);
        for (int l = 0; l < c.n_layers; l++) {
            const std::string p = "blk." + std::to_string(l) + ".";
            MegaLayerView L{};
            L.attn_norm = static_cast<const float*>(shard(*S, R, p + "attn_norm.weight", 0, r).ptr);
            L.ffn_norm = static_cast<const float*>(shard(*S, R, p + "ffn_norm.weight", 0, r).ptr);
            L.wq = shard(*S, R, p + "attn_q.weight", 1, r); L.wk = shard(*S, R, p + "attn_k.weight", 1, r);
            L.wv = shard(*S, R, p + "attn_v.weight", 1, r); L.wo = shard(*S, R, p + "attn_output.weight", 2, r);
            L.gate = shard(*S, R, p + "ffn_gate.weight", 1, r); L.up = shard(*S, R, p + "ffn_up.weight", 1, r);
            L.down = shard(*S, R, p + "ffn_down.weight", 2, r);
            R.kc[(size_t)l].alloc(kv_bytes); R.vc[(size_t)l].alloc(kv_bytes);
            L.kc = R.kc[(size_t)l].ptr; L.vc = R.vc[(size_t)l].ptr;
            mv.layers.push_back(L);
        }
        mv.head = shard(*S, R, S->file.find("output.weight") ? "output.weight" : "token_embd.weight", 1, r);
        mv.out_norm = static_cast<const float*>(shard(*S, R, "output_norm.weight", 0, r).ptr);
        R.logits.alloc((size_t)S->vocab_l * 4);
        mv.logits = R.logits.as<float>();
        R.step.alloc(8);
        mv.step = R.step.as<int>();
        const int qdim = nh * hd, kvdim = nkv * hd;
        R.hid[0].alloc((size_t)c.hidden_size * 4); R.hid[1].alloc((size_t)c.hidden_size * 4);
        R.q.alloc((size_t)qdim * 4); R.k.alloc((size_t)kvdim * 4); R.v.alloc((size_t)kvdim * 4); R.attn.alloc((size_t)qdim * 4);
        R.act.alloc((size_t)inter * 4);
        R.xq_h.alloc(xq_bytes(c.hidden_size)); R.xq_a.alloc(xq_bytes(qdim)); R.xq_i.alloc(xq_bytes(inter));
        R.sync.alloc(MEGA_SYNC_WORDS * 4);
        const size_t slot_floats = (size_t)2 * tp_size * c.hidden_size;
        R.xchg.alloc(slot_floats * 4 + (size_t)(tp_size + 1) * 32 * 4);
        MegaBuffers& B = R.B;
        B.hid[0] = R.hid[0].as<float>(); B.hid[1] = R.hid[1].as<float>(); B.q = R.q.as<float>(); B.k = R.k.as<float>();
        B.v = R.v.as<float>(); B.act = R.act.as<float>(); B.xq_h = R.xq_h.as<int8_t>(); B.xq_a = R.xq_a.as<int8_t>();
        B.xq_i = R.xq_i.as<int8_t>();
        std::string why;
        R.cnt_quant.alloc((size_t)inter / 32 * 4 + 4); R.cnt_attn.alloc((size_t)nh * 4 + 4);
        B.cnt_quant = R.cnt_quant.as<unsigned>(); B.cnt_attn = R.cnt_attn.as<unsigned>();
        R.cnt_norm.alloc((size_t)c.hidden_size / 32 * 4 + 4); R.ssq.alloc((size_t)c.hidden_size / 32 * 4 + 4);
        B.cnt_norm = R.cnt_norm.as<unsigned>(); B.ssq = R.ssq.as<float>();
        if (!mega_make_plan(mv, B, grid, split_fixed, fuse, &R.plan, &why)) { say("plan: " + why); return nullptr; }
        const std::string bad = mega_check_plan(R.plan, grid, tp_size);
        if (!bad.empty()) { say("plan check: " + bad); return nullptr; }
        R.scratch.alloc((size_t)nh * R.plan.n_splits_max * (hd + 2) * 4);
        MegaParams& P = R.P;
        memset(&P, 0, sizeof(P));
        P.phases = R.plan.phases.data(); P.first_gemv = R.plan.first_gemv;
        P.hidden = c.hidden_size; P.nh = nh; P.nkv = nkv; P.hd = hd; P.gc = R.plan.gc; P.max_seq = c.max_seq_len;
        P.eps = c.norm_eps; P.theta = c.rope_theta; P.freq_scale = c.rope_freq_scale; P.attn_scale = 1.0f / sqrtf((float)hd);
        P.step = R.step.as<int>();
        P.q = B.q; P.k = B.k; P.v = B.v; P.attn_out = R.attn.as<float>(); P.attn_scratch = R.scratch.as<float>(); P.xq_a = B.xq_a;
        P.n_splits_max = R.plan.n_splits_max; P.split_fixed = R.plan.split_fixed; P.min_split = R.plan.min_split; P.max_split = R.plan.max_split;
        P.attn_smem_off = R.plan.attn_smem_off;
        P.xchg_direct = (R.plan.fuse & MEGA_XCHG_DIRECT) ? 1 : 0;
        P.sync = R.sync.as<unsigned>();
        P.timeout_ns = 60ull * 1000000000ull;
        P.tp_rank = r; P.tp_size = tp_size;
    }
    // cross-link the ranks' slot / flag buffers (on hardware: cudaIpcOpenMemHandle over NVLink)
    const size_t slot_floats = (size_t)2 * tp_size * c.hidden_size;
    for (int r = 0; r < tp_size; r++)
        for (int o = 0; o < tp_size; o++) {
            S->ranks[(size_t)r]->P.slots[o] = S->ranks[(size_t)o]->xchg.as<float>();
            S->ranks[(size_t)r]->P.flags[o] = reinterpret_cast<unsigned*>(S->ranks[(size_t)o]->xchg.as<float>() + slot_floats);
        }
    say("ok");
    return S.release();
}

void mega_sim_free(void* h) { delete static_cast<Sim*>(h); }
int mega_sim_vocab(void* h) { return static_cast<Sim*>(h)->cfg.vocab_size; }

// One decode step on every rank concurrently.  embed_row: the token's embedding [hidden] (the product gathers it with a
// separate kernel).  logits_out: [vocab] assembled from the ranks' shards (null for prompt tokens whose logits are unused).
// Returns 0, 1 = emulator deadlock / time-out, 2 = a barrier inside the kernel timed out (abort word).
int mega_sim_step(void* h, const float* embed_row, int token, int pos, int with_head, float* logits_out) {
    Sim& S = *static_cast<Sim*>(h);
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    cusim::g_failed = false;
    for (int r = 0; r < S.tp; r++) {
        Rank& R = *S.ranks[(size_t)r];
        memcpy(R.hid[0].ptr, embed_row, (size_t)S.cfg.hidden_size * 4);
        R.step.as<int>()[0] = token; R.step.as<int>()[1] = pos;
        memset(R.sync.ptr, 0, 64 * 4);                                   // DecodeMega::launch's cudaMemsetAsync
        R.P.n_phases = with_head ? (int)R.plan.phases.size() : R.plan.n_body;
        if (const char* mp = getenv("NT_B200_MEGA_MAX_PHASES")) R.P.n_phases = std::max(1, std::min(R.P.n_phases, atoi(mp)));   // same bisect aid as DecodeMega::launch
    }
    // fault injection for the time-out test: one rank never launches (its peers must give up, not hang)
    const char* skip_env = getenv("CUSIM_MEGA_SKIP_RANK");
    const int skip_rank = skip_env ? atoi(skip_env) : -1;
    const char* to_env = getenv("CUSIM_MEGA_TIMEOUT_MS");
    for (int r = 0; r < S.tp; r++) {
        if (to_env) S.ranks[(size_t)r]->P.timeout_ns = (unsigned long long)atoll(to_env) * 1000000ull;
        if (r == skip_rank) continue;
        th.emplace_back([&, r]() {
            Rank& R = *S.ranks[(size_t)r];
            const MegaParams P = R.P;
            if (!cusim::launch(S.grid, NTHREADS, MEGA_DYN_SMEM, S.copy_delay, [P]() { decode_step_kernel(P); })) bad++;
        });
    }
    for (auto& t : th) t.join();
    if (bad.load() || cusim::g_failed.load()) return 1;
    for (int r = 0; r < S.tp; r++)
        if (S.ranks[(size_t)r]->sync.as<unsigned>()[64]) return 2;
    if (with_head && logits_out)
        for (int r = 0; r < S.tp; r++) {
            const int r0 = r * S.vocab_l, n = std::max(0, std::min(S.vocab_l, S.cfg.vocab_size - r0));
            memcpy(logits_out + r0, S.ranks[(size_t)r]->logits.ptr, (size_t)n * 4);
        }
    return 0;
}

// Debug: copy a working vector of rank r ("hid0", "hid1", "q", "k", "v", "attn", "act") to out; returns its length.
int mega_sim_read(void* h, int rank, const char* name, float* out, int cap) {
    Sim& S = *static_cast<Sim*>(h);
    Rank& R = *S.ranks[(size_t)rank];
    const std::string n = name;
    const MegaModelView& mv = R.mv;
    const float* p = nullptr;
    int len = 0;
    if (n == "hid0") { p = R.hid[0].as<float>(); len = mv.hidden; }
    else if (n == "hid1") { p = R.hid[1].as<float>(); len = mv.hidden; }
    else if (n == "q") { p = R.q.as<float>(); len = mv.nh * mv.hd; }
    else if (n == "k") { p = R.k.as<float>(); len = mv.nkv * mv.hd; }
    else if (n == "v") { p = R.v.as<float>(); len = mv.nkv * mv.hd; }
    else if (n == "attn") { p = R.attn.as<float>(); len = mv.nh * mv.hd; }
    else if (n == "act") { p = R.act.as<float>(); len = mv.inter; }
    else if (n == "scratch") { p = R.scratch.as<float>(); len = mv.nh * R.plan.n_splits_max * (mv.hd + 2); }
    else return -1;
    memcpy(out, p, (size_t)std::min(len, cap) * 4);
    return len;
}

}  // extern "C"
