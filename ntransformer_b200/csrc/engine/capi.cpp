// capi.cpp — extern "C" surface of the engine: include/ntransformer.h and include/nt_b200_engine.h.
#include "../../../include/ntransformer.h"
#include "../../../include/nt_b200_engine.h"
#include "engine.h"
#include "tp_comm.h"
#include <cstring>
#include <algorithm>
#include <memory>
#include <string>

using namespace nt::b200;

namespace {
struct ModelHandle {
    std::unique_ptr<TPComm> comm;     // declared first => destroyed last: the model's graphs reference the communicator
    Model model;
};
ModelHandle* H(nt_model_t m) { return static_cast<ModelHandle*>(m); }
}  // namespace

extern "C" {

// ---- include/ntransformer.h ----
nt_engine_t nt_engine_create(void) { return new Engine(); }
void nt_engine_destroy(nt_engine_t e) { delete static_cast<Engine*>(e); }
int nt_engine_load(nt_engine_t e, const char* path) { return (e && path && static_cast<Engine*>(e)->load(path)) ? 0 : -1; }
char* nt_engine_generate(nt_engine_t e, const char* prompt, int max_tokens, float temperature, int top_k, float top_p) {
    if (!e || !prompt) return nullptr;
    GenerateConfig cfg;
    cfg.max_tokens = max_tokens; cfg.temperature = temperature; cfg.top_k = top_k; cfg.top_p = top_p; cfg.verbose = false;
    std::string out = static_cast<Engine*>(e)->generate(prompt, cfg);
    char* r = static_cast<char*>(malloc(out.size() + 1));
    if (r) memcpy(r, out.c_str(), out.size() + 1);
    return r;
}
void nt_free(char* p) { free(p); }
int nt_engine_vocab_size(nt_engine_t e) { return e ? static_cast<Engine*>(e)->config().vocab_size : 0; }
int nt_engine_n_layers(nt_engine_t e) { return e ? static_cast<Engine*>(e)->config().n_layers : 0; }
int nt_engine_hidden_size(nt_engine_t e) { return e ? static_cast<Engine*>(e)->config().hidden_size : 0; }

// ---- include/nt_b200_engine.h ----
nt_model_t nt_model_load_gguf(const char* path, int max_context, int tp_rank, int tp_size) {
    auto* h = new ModelHandle();
    if (!path || !h->model.load_gguf(path, max_context, tp_rank, tp_size)) { delete h; return nullptr; }
    return h;
}
nt_model_t nt_model_create(const nt_model_config* c, int tp_rank, int tp_size) {
    if (!c) return nullptr;
    ModelConfig cfg;
    cfg.vocab_size = c->vocab_size; cfg.hidden_size = c->hidden_size; cfg.intermediate_size = c->intermediate_size;
    cfg.n_layers = c->n_layers; cfg.n_heads = c->n_heads; cfg.n_kv_heads = c->n_kv_heads; cfg.head_dim = c->head_dim;
    cfg.max_seq_len = c->max_seq_len; cfg.norm_eps = c->norm_eps; cfg.rope_theta = c->rope_theta;
    cfg.bos_token_id = c->bos_token_id; cfg.eos_token_id = c->eos_token_id;
    auto* h = new ModelHandle();
    h->model.init(cfg, tp_rank, tp_size);
    return h;
}
int nt_model_set_tensor(nt_model_t m, const char* name, const void* p, int dtype, size_t pitch) {
    return (m && name && H(m)->model.set_tensor(name, p, (nt::DType)dtype, pitch)) ? 0 : -1;
}
int nt_model_finalize(nt_model_t m) { return (m && H(m)->model.finalize()) ? 0 : -1; }
void nt_model_free(nt_model_t m) { delete H(m); }
int nt_model_get_config(nt_model_t m, nt_model_config* o) {
    if (!m || !o) return -1;
    const ModelConfig& c = H(m)->model.config();
    *o = nt_model_config{c.vocab_size, c.hidden_size, c.intermediate_size, c.n_layers, c.n_heads, c.n_kv_heads, c.head_dim,
                         c.max_seq_len, c.norm_eps, c.rope_theta, c.bos_token_id, c.eos_token_id};
    return 0;
}
int nt_model_forward(nt_model_t m, const int* tokens, int n, int start_pos, float* logits_host) {
    if (!m || !tokens) return -1;
    float* dl = H(m)->model.forward(tokens, n, start_pos);
    if (logits_host)
        NT_CUDA_CHECK(cudaMemcpy(logits_host, dl, sizeof(float) * (size_t)H(m)->model.config().vocab_size, cudaMemcpyDeviceToHost));
    return 0;
}
int nt_model_forward_async(nt_model_t m, const int* tokens, int n, int start_pos) {
    if (!m || !tokens) return -1;
    H(m)->model.forward_async(tokens, n, start_pos);
    return 0;
}
int nt_model_sync(nt_model_t m) { return m ? H(m)->model.sync() : -1; }
float* nt_model_logits_device(nt_model_t m) { return m ? H(m)->model.logits_device() : nullptr; }
void* nt_model_stream(nt_model_t m) { return m ? (void*)H(m)->model.stream() : nullptr; }
int nt_model_argmax(nt_model_t m) { return m ? H(m)->model.argmax_last() : -1; }
void nt_model_clear_kv(nt_model_t m) { if (m) H(m)->model.clear_kv(); }
void nt_model_set_prefill_min_tokens(nt_model_t m, int n) { if (m) H(m)->model.set_prefill_min_tokens(n); }
void nt_model_use_graph(nt_model_t m, int on) { if (m) H(m)->model.set_use_graph(on != 0); }
unsigned long long nt_model_bytes_per_token(nt_model_t m, int ctx) { return m ? H(m)->model.bytes_per_token(ctx) : 0; }
int nt_model_sample(nt_model_t m, float temperature, int top_k, float top_p, float repeat_penalty, const int* recent_window, int n_window,
                    float r) {
    return m ? H(m)->model.sample_last(temperature, top_k, top_p, repeat_penalty, recent_window, n_window, r) : -1;
}
float nt_sampler_uniform(uint64_t seed, int n_draws_before) {
    SamplerConfig sc;
    sc.seed = seed;
    Sampler s;
    s.init(sc);
    for (int i = 0; i < n_draws_before; i++) s.draw();
    return s.draw();
}
void nt_model_use_megakernel(nt_model_t m, int on) { if (m) H(m)->model.set_use_megakernel(on != 0); }
double nt_model_load_seconds(nt_model_t m) { return m ? H(m)->model.load_seconds() : 0.0; }
int nt_model_tp_exchange(nt_model_t m) { return m ? H(m)->model.tp_exchange_kind() : 0; }
int nt_model_megakernel_active(nt_model_t m) { return (m && H(m)->model.megakernel_active()) ? 1 : 0; }
int nt_model_megakernel_plan(nt_model_t m, int* kinds, int cap) {
    if (!m) return -1;
    return H(m)->model.mega_plan_kinds(kinds, cap);
}
void nt_model_megakernel_trace(nt_model_t m, int on) { if (m) H(m)->model.mega_trace(on != 0); }
long long nt_model_megakernel_trace_read(nt_model_t m, unsigned long long* out_host, size_t cap) {
    return m ? (long long)H(m)->model.mega_trace_read(out_host, cap) : -1;
}
long long nt_model_debug_read(nt_model_t m, const char* name, float* out_host, size_t cap) {
    if (!m || !name) return -1;
    size_t n = 0;
    const float* p = H(m)->model.mega_debug_buffer(name, &n);
    if (!p) return -1;
    NT_CUDA_CHECK(cudaStreamSynchronize(H(m)->model.stream()));
    if (out_host) NT_CUDA_CHECK(cudaMemcpy(out_host, p, sizeof(float) * (n < cap ? n : cap), cudaMemcpyDeviceToHost));
    return (long long)n;
}

// Host-only: builds the persistent kernel's per-token program for a model of the given (per-rank) shape with placeholder
// addresses and replays its schedule on the CPU (decode_mega.h: mega_make_plan + mega_check_plan).  No CUDA call.
int nt_mega_plan_selftest(const nt_model_config* c, int tp_rank, int tp_size, const int* layer_dtypes, int head_dtype, int grid,
                          int split_fixed, int fuse, int* info, char* msg, size_t cap) {
    auto say = [&](const std::string& m) { if (msg && cap) { snprintf(msg, cap, "%s", m.c_str()); } };
    if (!c || !layer_dtypes || tp_size < 1 || c->n_heads % tp_size || c->n_kv_heads % tp_size || c->intermediate_size % tp_size) {
        say("bad arguments");
        return -1;
    }
    MegaModelView mv;
    mv.hidden = c->hidden_size; mv.nh = c->n_heads / tp_size; mv.nkv = c->n_kv_heads / tp_size; mv.hd = c->head_dim;
    mv.inter = c->intermediate_size / tp_size; mv.max_seq = c->max_seq_len; mv.n_layers = c->n_layers;
    mv.eps = c->norm_eps; mv.theta = c->rope_theta; mv.tp_rank = tp_rank; mv.tp_size = tp_size;
    uintptr_t next = 0x10000000;                       // placeholder device addresses, 256-byte aligned
    auto place = [&](size_t bytes) { void* p = reinterpret_cast<void*>(next); next += (bytes + 255) & ~(size_t)255; return p; };
    auto weight = [&](int dt, int rows, int cols, bool col_shard) {
        MegaWeight w;
        w.dtype = (nt::DType)dt; w.rows = rows; w.cols = cols;
        const size_t rb = nt::dtype_row_size(w.dtype, (size_t)cols);
        w.pitch = col_shard ? ((rb + 15) & ~(size_t)15) : rb;      // Model::upload pads column shards to a 16-byte pitch
        w.ptr = place(w.pitch * (size_t)std::max(rows, 1));
        return w;
    };
    const int qdim = mv.nh * mv.hd, kvdim = mv.nkv * mv.hd;
    for (int l = 0; l < mv.n_layers; l++) {
        const int* d = layer_dtypes + 7 * l;
        MegaLayerView L;
        L.attn_norm = static_cast<const float*>(place((size_t)mv.hidden * 4));
        L.ffn_norm = static_cast<const float*>(place((size_t)mv.hidden * 4));
        L.wq = weight(d[0], qdim, mv.hidden, false); L.wk = weight(d[1], kvdim, mv.hidden, false); L.wv = weight(d[2], kvdim, mv.hidden, false);
        L.wo = weight(d[3], mv.hidden, qdim, tp_size > 1);
        L.gate = weight(d[4], mv.inter, mv.hidden, false); L.up = weight(d[5], mv.inter, mv.hidden, false);
        L.down = weight(d[6], mv.hidden, mv.inter, tp_size > 1);
        L.kc = place((size_t)mv.max_seq * kvdim * 2); L.vc = place((size_t)mv.max_seq * kvdim * 2);
        mv.layers.push_back(L);
    }
    const int vl = (c->vocab_size + tp_size - 1) / tp_size;
    const int vrows = tp_size == 1 ? c->vocab_size : std::max(0, std::min(vl, c->vocab_size - tp_rank * vl));
    mv.head = weight(head_dtype, vrows, mv.hidden, false);
    mv.out_norm = static_cast<const float*>(place((size_t)mv.hidden * 4));
    mv.logits = static_cast<float*>(place((size_t)vl * 4));
    mv.step = static_cast<const int*>(place(8));
    MegaBuffers B;
    B.hid[0] = static_cast<float*>(place((size_t)mv.hidden * 4)); B.hid[1] = static_cast<float*>(place((size_t)mv.hidden * 4));
    B.q = static_cast<float*>(place((size_t)qdim * 4)); B.k = static_cast<float*>(place((size_t)kvdim * 4));
    B.v = static_cast<float*>(place((size_t)kvdim * 4)); B.act = static_cast<float*>(place((size_t)mv.inter * 4));
    B.xq_h = static_cast<int8_t*>(place(xq_bytes(mv.hidden))); B.xq_a = static_cast<int8_t*>(place(xq_bytes(qdim)));
    B.xq_i = static_cast<int8_t*>(place(xq_bytes(mv.inter)));
    B.cnt_quant = static_cast<unsigned*>(place((size_t)mv.inter / 32 * 4 + 4)); B.cnt_attn = static_cast<unsigned*>(place((size_t)mv.nh * 4 + 4));
    B.cnt_norm = static_cast<unsigned*>(place((size_t)mv.hidden / 32 * 4 + 4)); B.ssq = static_cast<float*>(place((size_t)mv.hidden / 32 * 4 + 4));
    MegaPlan pl;
    std::string why;
    if (!mega_make_plan(mv, B, grid, split_fixed, fuse, &pl, &why)) { say(why); return 1; }
    if (info) {
        int n_gemv = 0, min_warps = 99, min_stages = 99, n_exchange = 0;
        for (const MegaPhase& ph : pl.phases) {
            if (ph.kind == MPH_GEMV) { n_gemv++; min_warps = std::min(min_warps, ph.warps); min_stages = std::min(min_stages, ph.stages); }
            if (ph.barrier == MBAR_EXCHANGE) n_exchange++;
        }
        info[0] = (int)pl.phases.size(); info[1] = pl.n_body; info[2] = n_gemv; info[3] = min_warps; info[4] = min_stages;
        info[5] = n_exchange; info[6] = pl.n_splits_max; info[7] = pl.max_split;
    }
    const std::string bad = mega_check_plan(pl, grid, tp_size);
    if (!bad.empty()) { say(bad); return 2; }
    say("ok");
    return 0;
}

int nt_gguf_describe(const char* path, char* out, size_t cap) {
    if (!path || !out || !cap) return -1;
    GGUFFile f;
    if (!f.open(path)) return -1;
    const ModelConfig& c = f.config();
    std::string j = "{";
    auto kv = [&](const char* k, long long v) { j += "\"" + std::string(k) + "\": " + std::to_string(v) + ", "; };
    kv("vocab_size", c.vocab_size); kv("hidden_size", c.hidden_size); kv("intermediate_size", c.intermediate_size);
    kv("n_layers", c.n_layers); kv("n_heads", c.n_heads); kv("n_kv_heads", c.n_kv_heads); kv("head_dim", c.head_dim);
    kv("max_seq_len", c.max_seq_len); kv("bos_token_id", c.bos_token_id); kv("eos_token_id", c.eos_token_id);
    kv("n_vocab_tokens", (long long)f.vocab().tokens.size()); kv("data_offset", (long long)f.data_offset());
    char fl[96];
    snprintf(fl, sizeof(fl), "\"norm_eps\": %.9g, \"rope_theta\": %.9g, ", c.norm_eps, c.rope_theta);
    j += fl;
    j += "\"architecture\": \"" + c.architecture + "\", \"tensors\": [";
    bool first = true;
    for (const auto& t : f.tensors()) {
        if (!first) j += ", ";
        first = false;
        j += "{\"name\": \"" + t.name + "\", \"dtype\": " + std::to_string((int)t.dtype) + ", \"offset\": " + std::to_string(t.offset) +
             ", \"nbytes\": " + std::to_string(t.nbytes) + ", \"shape\": [";
        for (size_t d = 0; d < t.shape.size(); d++) j += (d ? ", " : "") + std::to_string(t.shape[d]);
        j += "]}";
    }
    j += "]}";
    if (j.size() + 1 > cap) return -1;
    memcpy(out, j.c_str(), j.size() + 1);
    return (int)j.size();
}
int nt_tokenize(const char* gguf_path, const char* text, int add_bos, int* ids, int cap) {
    if (!gguf_path || !text) return -1;
    GGUFFile f;
    if (!f.open(gguf_path)) return -1;
    Tokenizer t;
    t.init(f.vocab(), f.config().bos_token_id, f.config().eos_token_id);
    std::vector<int> v = t.encode(text, add_bos != 0);
    for (int i = 0; i < (int)v.size() && i < cap; i++) ids[i] = v[(size_t)i];
    return (int)v.size();
}
int nt_detokenize(const char* gguf_path, const int* ids, int n, char* out, size_t cap) {
    if (!gguf_path || !out || !cap) return -1;
    GGUFFile f;
    if (!f.open(gguf_path)) return -1;
    Tokenizer t;
    t.init(f.vocab(), f.config().bos_token_id, f.config().eos_token_id);
    std::string s = t.decode(std::vector<int>(ids, ids + n));
    if (s.size() + 1 > cap) return -1;
    memcpy(out, s.data(), s.size());
    out[s.size()] = 0;
    return (int)s.size();
}
int nt_sample_token(const float* logits, int n, float temperature, int top_k, float top_p, float repeat_penalty,
                    int repeat_window, const int* recent, int n_recent, uint64_t seed) {
    if (!logits || n <= 0) return -1;
    SamplerConfig sc;
    sc.temperature = temperature; sc.top_k = top_k; sc.top_p = top_p; sc.repeat_penalty = repeat_penalty;
    sc.repeat_window = repeat_window; sc.seed = seed;
    Sampler s;
    s.init(sc);
    std::vector<float> l(logits, logits + n);
    s.apply_repeat_penalty(l.data(), n, std::vector<int>(recent, recent + (recent ? n_recent : 0)));
    return s.sample(l.data(), n);
}

int nt_tp_unique_id(void* out128) { return TPComm::unique_id(out128) ? 0 : -1; }
int nt_tp_init(nt_model_t m, const void* id128, int rank, int size) {
    if (!m) return -1;
    auto c = std::make_unique<TPComm>();
    if (!c->init(id128, rank, size)) return -2;
    {   // NCCL builds its channels lazily on the first collective (allocations, IPC handles): do that here, eagerly,
        // so that nothing of the sort happens inside the decode step's CUDA-graph capture.
        cudaStream_t st = H(m)->model.stream();
        float* tmp = nullptr;
        NT_CUDA_CHECK(cudaMalloc(&tmp, sizeof(float) * 1024 * (size_t)(size + 1)));
        NT_CUDA_CHECK(cudaMemsetAsync(tmp, 0, sizeof(float) * 1024 * (size_t)(size + 1), st));
        c->all_reduce_sum(tmp, 1024, st);
        c->all_gather(tmp, tmp + 1024, 1024, st);
        NT_CUDA_CHECK(cudaStreamSynchronize(st));
        NT_CUDA_CHECK(cudaFree(tmp));
    }
    H(m)->model.set_comm(c.get());
    H(m)->comm = std::move(c);
    return 0;
}
}  // extern "C"
