// model.cu — see model.h.
#include "model.h"
#include "tp_comm.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace nt { namespace b200 {

namespace {

__global__ void set_step_kernel(int* step, int token, int pos) { step[0] = token; step[1] = pos; }
__global__ void iota_kernel(int* out, int first, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first + i;
}
constexpr int PREFILL_CHUNK = 4096;      // tokens per batched pass (bounds the activation buffers; weights re-read per chunk)

// argmax over n floats, lowest index wins ties (reference Sampler::argmax, sampler.cpp:18-28)
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ x, int n, int* __restrict__ out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float best = -FLT_MAX;
    int bi = 0x7FFFFFFF;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = x[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = sv[threadIdx.x]; bi = si[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) *out = bi;
    }
}

template <typename T> T* dmalloc(size_t n) {
    T* p = nullptr;
    NT_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}
size_t round16(size_t v) { return (v + 15) & ~(size_t)15; }
bool env_on(const char* name) { const char* e = getenv(name); return e && *e && strcmp(e, "0") != 0; }   // set and not "0"

}  // namespace

Model::~Model() {
    if (stream_) cudaStreamSynchronize(stream_);
    release_graphs();
    for (void* p : owned_) cudaFree(p);
    for (void* p : {(void*)hidden_, (void*)xnorm_, (void*)q_, (void*)k_, (void*)v_, (void*)attn_, (void*)act_, (void*)up_, (void*)part_,
                    (void*)logits_, (void*)logits_l_, (void*)attn_scratch_, xq_h_, xq_a_, xq_i_, kc_, vc_, (void*)step_dev_,
                    (void*)argmax_dev_, (void*)recent_dev_})
        if (p) cudaFree(p);
    for (void* p : {(void*)pf_.x, (void*)pf_.q, (void*)pf_.k, (void*)pf_.v, (void*)pf_.attn, pf_.ws, pf_.ws2, (void*)pf_.tok, (void*)pf_.pos,
                    (void*)pf_.g, (void*)pf_.u, pf_.whi, pf_.wlo})
        if (p) cudaFree(p);
    if (argmax_host_) cudaFreeHost(argmax_host_);
    if (stream_) cudaStreamDestroy(stream_);
}

void Model::release_graphs() {
    if (g_full_) { cudaGraphExecDestroy(g_full_); g_full_ = nullptr; }
    if (g_body_) { cudaGraphExecDestroy(g_body_); g_body_ = nullptr; }
}

void Model::init(const ModelConfig& cfg, int tp_rank, int tp_size) {
    cfg_ = cfg;
    tp_rank_ = tp_rank;
    tp_size_ = tp_size;
    NT_CHECK(tp_size >= 1 && tp_rank >= 0 && tp_rank < tp_size, "bad tensor-parallel rank/size");
    NT_CHECK(cfg.n_heads % tp_size == 0 && cfg.n_kv_heads % tp_size == 0, "heads must divide by tp_size");
    NT_CHECK(cfg.intermediate_size % tp_size == 0, "intermediate_size must divide by tp_size");
    nh_l_ = cfg.n_heads / tp_size;
    nkv_l_ = cfg.n_kv_heads / tp_size;
    inter_l_ = cfg.intermediate_size / tp_size;
    vocab_l_ = (cfg.vocab_size + tp_size - 1) / tp_size;
    layers_.assign((size_t)cfg.n_layers, LayerWeights{});
}

bool Model::set_tensor(const std::string& name, const void* ptr, DType dt, size_t pitch) {
    const int hidden = cfg_.hidden_size, hd = cfg_.head_dim;
    auto W = [&](Weight& w, int rows, int cols) {
        w.ptr = ptr; w.dtype = dt; w.rows = rows; w.cols = cols;
        w.pitch = pitch ? pitch : dtype_row_size(dt, (size_t)cols);
        w.owned = false;
        return true;
    };
    if (name == "token_embd.weight") return W(embd_, cfg_.vocab_size, hidden);
    if (name == "output.weight") {
        int rows = (tp_size_ == 1) ? cfg_.vocab_size : std::max(0, std::min(vocab_l_, cfg_.vocab_size - tp_rank_ * vocab_l_));
        return W(head_, rows, hidden);
    }
    if (name == "output_norm.weight") { out_norm_ = static_cast<const float*>(ptr); return true; }
    int li = -1;
    char field[64] = {0};
    if (sscanf(name.c_str(), "blk.%d.%63s", &li, field) != 2 || li < 0 || li >= cfg_.n_layers) return false;
    LayerWeights& L = layers_[(size_t)li];
    const std::string f = field;
    if (f == "attn_norm.weight") { L.attn_norm = static_cast<const float*>(ptr); return true; }
    if (f == "ffn_norm.weight") { L.ffn_norm = static_cast<const float*>(ptr); return true; }
    if (f == "attn_q.weight") return W(L.wq, nh_l_ * hd, hidden);
    if (f == "attn_k.weight") return W(L.wk, nkv_l_ * hd, hidden);
    if (f == "attn_v.weight") return W(L.wv, nkv_l_ * hd, hidden);
    if (f == "attn_output.weight") return W(L.wo, hidden, nh_l_ * hd);
    if (f == "ffn_gate.weight") return W(L.gate, inter_l_, hidden);
    if (f == "ffn_up.weight") return W(L.up, inter_l_, hidden);
    if (f == "ffn_down.weight") return W(L.down, hidden, inter_l_);
    return false;
}

// Uploads tensor `name` (sharded) and fills *w. split: 0 replicate, 1 rows, 2 columns (quant-block aligned).
const void* Model::upload(const GGUFFile& f, const std::string& name, Weight* w, int split) {
    const GGUFTensorInfo* ti = f.find(name);
    NT_CHECK(ti != nullptr, ("Tensor not found: " + name).c_str());
    const uint8_t* src = static_cast<const uint8_t*>(f.data(*ti));
    const DType dt = ti->dtype;
    const int cols = (int)ti->shape[0];
    const int rows = ti->shape.size() > 1 ? (int)ti->shape[1] : 1;
    const size_t row_bytes = dtype_row_size(dt, (size_t)cols);
    void* dst = nullptr;
    if (split == 0 || tp_size_ == 1) {
        NT_CUDA_CHECK(cudaMalloc(&dst, std::max<size_t>(ti->nbytes, 16)));
        NT_CUDA_CHECK(cudaMemcpy(dst, src, ti->nbytes, cudaMemcpyHostToDevice));
        if (w) { w->rows = rows; w->cols = cols; w->pitch = row_bytes; }
    } else if (split == 1) {
        int per = (rows + tp_size_ - 1) / tp_size_;
        int r0 = std::min(rows, tp_rank_ * per), r1 = std::min(rows, r0 + per);
        size_t bytes = (size_t)(r1 - r0) * row_bytes;
        NT_CUDA_CHECK(cudaMalloc(&dst, std::max<size_t>(bytes, 16)));
        if (bytes) NT_CUDA_CHECK(cudaMemcpy(dst, src + (size_t)r0 * row_bytes, bytes, cudaMemcpyHostToDevice));
        if (w) { w->rows = r1 - r0; w->cols = cols; w->pitch = row_bytes; }
    } else {
        const int bs = (int)dtype_block_size(dt);
        NT_CHECK(cols % tp_size_ == 0 && (cols / tp_size_) % bs == 0, "column shard is not quant-block aligned");
        const int c_l = cols / tp_size_;
        const size_t shard_row = dtype_row_size(dt, (size_t)c_l);
        const size_t pitch = round16(shard_row);                // rows start 16 B aligned for the TMA path
        NT_CUDA_CHECK(cudaMalloc(&dst, std::max<size_t>(pitch * rows, 16)));
        NT_CUDA_CHECK(cudaMemset(dst, 0, pitch * rows));
        NT_CUDA_CHECK(cudaMemcpy2D(dst, pitch, src + (size_t)tp_rank_ * shard_row, row_bytes, shard_row, (size_t)rows,
                                   cudaMemcpyHostToDevice));
        if (w) { w->rows = rows; w->cols = c_l; w->pitch = pitch; }
    }
    owned_.push_back(dst);
    if (w) { w->ptr = dst; w->dtype = dt; w->owned = true; }
    return dst;
}

bool Model::load_gguf(const std::string& path, int max_context, int tp_rank, int tp_size) {
    fprintf(stderr, "Loading model: %s\n", path.c_str());
    GGUFFile f;
    if (!f.open(path)) return false;
    ModelConfig cfg = f.config();
    if (cfg.max_seq_len > max_context) {          // transformer.cpp:70-74
        fprintf(stderr, "Note: Capping context from %d to %d tokens (use --ctx-size to change)\n", cfg.max_seq_len, max_context);
        cfg.max_seq_len = max_context;
    }
    cfg.print();
    f.print_info();
    vocab_ = f.vocab();
    init(cfg, tp_rank, tp_size);

    upload(f, "token_embd.weight", &embd_, 0);
    if (embd_.dtype == DType::Q5_K)
        fprintf(stderr, "Error: Unsupported embedding dtype: Q5_K (rows read as zeros, like the reference)\n");
    if (f.find("output.weight")) {
        upload(f, "output.weight", &head_, 1);
    } else if (tp_size_ == 1) {
        head_ = embd_;                               // tied embeddings (transformer.cpp:96-99)
        head_.owned = false;
    } else {
        upload(f, "token_embd.weight", &head_, 1);
    }
    out_norm_ = static_cast<const float*>(upload(f, "output_norm.weight", nullptr, 0));
    for (int i = 0; i < cfg_.n_layers; i++) {
        const std::string p = "blk." + std::to_string(i) + ".";
        LayerWeights& L = layers_[(size_t)i];
        L.attn_norm = static_cast<const float*>(upload(f, p + "attn_norm.weight", nullptr, 0));
        upload(f, p + "attn_q.weight", &L.wq, 1);
        upload(f, p + "attn_k.weight", &L.wk, 1);
        upload(f, p + "attn_v.weight", &L.wv, 1);
        upload(f, p + "attn_output.weight", &L.wo, 2);
        L.ffn_norm = static_cast<const float*>(upload(f, p + "ffn_norm.weight", nullptr, 0));
        upload(f, p + "ffn_gate.weight", &L.gate, 1);
        upload(f, p + "ffn_up.weight", &L.up, 1);
        upload(f, p + "ffn_down.weight", &L.down, 2);
        if ((i & 7) == 7 || i + 1 == cfg_.n_layers) fprintf(stderr, "  Loaded layer %d/%d\n", i + 1, cfg_.n_layers);
    }
    if (!finalize()) return false;
    fprintf(stderr, "Model loaded successfully!\n");
    return true;
}

bool Model::finalize() {
    const int hidden = cfg_.hidden_size, hd = cfg_.head_dim;
    if (!embd_.ptr || !out_norm_) { fprintf(stderr, "model: token_embd / output_norm missing\n"); return false; }
    if (!head_.ptr) {
        if (tp_size_ != 1) { fprintf(stderr, "model: output.weight shard missing\n"); return false; }
        head_ = embd_; head_.owned = false;
    }
    for (int i = 0; i < cfg_.n_layers; i++) {
        const LayerWeights& L = layers_[(size_t)i];
        if (!L.attn_norm || !L.ffn_norm || !L.wq.ptr || !L.wk.ptr || !L.wv.ptr || !L.wo.ptr || !L.gate.ptr || !L.up.ptr || !L.down.ptr) {
            fprintf(stderr, "model: layer %d is missing tensors\n", i);
            return false;
        }
    }
    NT_CHECK(hidden % 32 == 0 && (nh_l_ * hd) % 32 == 0 && inter_l_ % 32 == 0, "dimensions must be multiples of 32");
    NT_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    const int max_seq = cfg_.max_seq_len;
    hidden_ = dmalloc<float>(hidden);
    xnorm_ = dmalloc<float>(hidden);
    q_ = dmalloc<float>((size_t)nh_l_ * hd);
    k_ = dmalloc<float>((size_t)nkv_l_ * hd);
    v_ = dmalloc<float>((size_t)nkv_l_ * hd);
    attn_ = dmalloc<float>((size_t)nh_l_ * hd);
    act_ = dmalloc<float>(inter_l_);
    up_ = dmalloc<float>(inter_l_);
    part_ = dmalloc<float>(hidden);
    logits_ = dmalloc<float>((size_t)vocab_l_ * tp_size_);
    logits_l_ = dmalloc<float>(vocab_l_);
    attn_scratch_ = dmalloc<float>(attention_decode_dyn_scratch_floats(max_seq, nh_l_, nkv_l_, hd));
    xq_h_ = dmalloc<uint8_t>(xq_bytes(hidden));
    xq_a_ = dmalloc<uint8_t>(xq_bytes(nh_l_ * hd));
    xq_i_ = dmalloc<uint8_t>(xq_bytes(inter_l_));
    const size_t kv_elems = (size_t)cfg_.n_layers * max_seq * nkv_l_ * hd;     // F16 [L, max_seq, n_kv, hd], transformer.cpp:340-346
    kc_ = dmalloc<uint16_t>(kv_elems);
    vc_ = dmalloc<uint16_t>(kv_elems);
    NT_CUDA_CHECK(cudaMemset(kc_, 0, kv_elems * 2));
    NT_CUDA_CHECK(cudaMemset(vc_, 0, kv_elems * 2));
    NT_CUDA_CHECK(cudaMemset(logits_, 0, sizeof(float) * (size_t)vocab_l_ * tp_size_));
    step_dev_ = dmalloc<int>(2);
    argmax_dev_ = dmalloc<int>(1);
    NT_CUDA_CHECK(cudaMallocHost(&argmax_host_, sizeof(int)));
    finalized_ = true;
    return true;
}

void Model::clear_kv() {
    const size_t kv_elems = (size_t)cfg_.n_layers * cfg_.max_seq_len * nkv_l_ * cfg_.head_dim;
    NT_CUDA_CHECK(cudaMemsetAsync(kc_, 0, kv_elems * 2, stream_));
    NT_CUDA_CHECK(cudaMemsetAsync(vc_, 0, kv_elems * 2, stream_));
}

size_t Model::weight_bytes() const {
    size_t b = head_.bytes() + (size_t)(2 * cfg_.n_layers + 1) * cfg_.hidden_size * 4;
    for (const LayerWeights& L : layers_) b += L.wq.bytes() + L.wk.bytes() + L.wv.bytes() + L.wo.bytes() + L.gate.bytes() + L.up.bytes() + L.down.bytes();
    return b;
}
size_t Model::bytes_per_token(int ctx) const {
    return weight_bytes() + (size_t)2 * cfg_.n_layers * ctx * nkv_l_ * cfg_.head_dim * 2;
}

// y_i = W_i . x' for n matrices sharing x' = norm_w ? rmsnorm(x, norm_w) : x.
// When every matrix qualifies, one fused launch does (RMSNorm +) activation quantisation in its prologue, the TMA/dp4a
// GEMV and the epilogue; otherwise the pieces run as separate launches like the reference's sequence.
void Model::matvec(const Weight* const* ws, float* const* ys, int n, const float* x, const float* norm_w, GemvEpilogue ep,
                   cudaStream_t s) {
    GemvMat mats[3];
    const int K = ws[0]->cols;
    for (int i = 0; i < n; i++) { mats[i].W = ws[i]->ptr; mats[i].y = ys[i]; mats[i].out = ws[i]->rows; mats[i].dtype = ws[i]->dtype; mats[i].row_pitch = ws[i]->pitch; }
    GemvInput in;
    if (gemv_kq_supported(mats, n, K)) {
        // (RMSNorm +) quantisation as one small multi-CTA launch: measured cheaper than redoing it in the prologue
        // of each of the 148 GEMV CTAs (profiles/r01_*), and PDL overlaps it with the neighbouring kernels.
        void* xq = (K == cfg_.hidden_size) ? xq_h_ : (K == inter_l_) ? xq_i_ : xq_a_;
        if (norm_w) rmsnorm_xq(nullptr, xq, x, norm_w, K, cfg_.norm_eps, s);
        else quantize_x(x, xq, K, s);
        in.xq = xq;
        gemv_kq(mats, n, K, in, ep, s);
        return;
    }
    const float* xf = x;
    if (norm_w) { rmsnorm(xnorm_, x, norm_w, 1, K, cfg_.norm_eps, s); xf = xnorm_; }
    if (ep == GEMV_SWIGLU) {                   // unfused fallback: gate -> ys[0], up -> ys[1], then silu_mul in place
        const Weight* g = ws[0]; const Weight* u = ws[1];
        matvec(&g, &ys[0], 1, xf, nullptr, GEMV_STORE, s);
        matvec(&u, &ys[1], 1, xf, nullptr, GEMV_STORE, s);
        silu_mul(ys[0], ys[0], ys[1], ws[0]->rows, s);
        return;
    }
    for (int i = 0; i < n; i++) {
        if (gemv_kq_supported(&mats[i], 1, K)) { in.x = xf; gemv_kq(&mats[i], 1, K, in, ep, s); }
        else gemv_generic(ys[i], ws[i]->ptr, xf, ws[i]->rows, K, ws[i]->dtype, ws[i]->pitch, ep, s);
    }
}

bool Model::o_xq_fusable(const Weight& wo) const {
    GemvMat m; m.W = wo.ptr; m.y = nullptr; m.out = wo.rows; m.dtype = wo.dtype; m.row_pitch = wo.pitch;
    return (nh_l_ * cfg_.head_dim) % 128 == 0 && gemv_kq_supported(&m, 1, wo.cols);
}

// hidden += all_reduce(partial)   (one exchange per sub-block)
void Model::reduce_residual(float* partial, cudaStream_t s) {
    NT_CHECK(comm_ != nullptr, "tensor-parallel model without a communicator");
    comm_->all_reduce_sum(partial, (size_t)cfg_.hidden_size, s);
    add_inplace(hidden_, partial, cfg_.hidden_size, s);
}

void Model::step_body(cudaStream_t s) {
    const int hidden = cfg_.hidden_size, hd = cfg_.head_dim, max_seq = cfg_.max_seq_len;
    const float scale = 1.0f / sqrtf((float)hd);                         // attention.cpp:21
    const size_t kv_stride = (size_t)max_seq * nkv_l_ * hd;               // transformer.cpp:629
    const int* pos_dev = step_dev_ + 1;
    embed_rows(hidden_, embd_.ptr, embd_.dtype, step_dev_, 1, hidden, s);
    for (int i = 0; i < cfg_.n_layers; i++) {
        const LayerWeights& L = layers_[(size_t)i];
        uint16_t* kc = static_cast<uint16_t*>(kc_) + (size_t)i * kv_stride;
        uint16_t* vc = static_cast<uint16_t*>(vc_) + (size_t)i * kv_stride;
        // --- attention sub-block ---
        { const Weight* ws[3] = {&L.wq, &L.wk, &L.wv}; float* ys[3] = {q_, k_, v_}; matvec(ws, ys, 3, hidden_, L.attn_norm, GEMV_STORE, s); }
        rope_kv_decode(q_, k_, v_, kc, vc, pos_dev, nh_l_, nkv_l_, hd, cfg_.rope_theta, cfg_.rope_freq_scale, max_seq, s);
        const bool o_fused = o_xq_fusable(L.wo);          // attention emits xq for the o-projection directly
        attention_decode_dyn(attn_, q_, kc, vc, pos_dev, max_seq, nh_l_, nkv_l_, hd, scale, attn_scratch_, o_fused ? xq_a_ : nullptr, s);
        { const Weight* ws[1] = {&L.wo};
          float* ys[1] = {tp_size_ == 1 ? hidden_ : part_};
          const GemvEpilogue ep = tp_size_ == 1 ? GEMV_ADD : GEMV_STORE;
          if (o_fused) {
              GemvMat m; m.W = L.wo.ptr; m.y = ys[0]; m.out = L.wo.rows; m.dtype = L.wo.dtype; m.row_pitch = L.wo.pitch;
              gemv_kq(&m, 1, L.wo.cols, xq_a_, ep, s);
          } else {
              matvec(ws, ys, 1, attn_, nullptr, ep, s);
          }
          if (tp_size_ > 1) reduce_residual(part_, s); }
        // --- FFN sub-block ---
        { const Weight* ws[2] = {&L.gate, &L.up}; float* ys[2] = {act_, up_}; matvec(ws, ys, 2, hidden_, L.ffn_norm, GEMV_SWIGLU, s); }
        { const Weight* ws[1] = {&L.down};
          if (tp_size_ == 1) { float* ys[1] = {hidden_}; matvec(ws, ys, 1, act_, nullptr, GEMV_ADD, s); }
          else { float* ys[1] = {part_}; matvec(ws, ys, 1, act_, nullptr, GEMV_STORE, s); reduce_residual(part_, s); } }
    }
}

void Model::step_head(cudaStream_t s) {
    const Weight* ws[1] = {&head_};                                               // final norm fused: transformer.cpp:657-665
    if (tp_size_ == 1) {
        float* ys[1] = {logits_};
        matvec(ws, ys, 1, hidden_, out_norm_, GEMV_STORE, s);
    } else {
        float* ys[1] = {logits_l_};
        if (head_.rows > 0) matvec(ws, ys, 1, hidden_, out_norm_, GEMV_STORE, s);
        comm_->all_gather(logits_l_, logits_, (size_t)vocab_l_, s);
    }
}

// ---- opt-in: the decode step as one persistent kernel (engine/decode_mega.h) ----
bool Model::ensure_mega() {
    if (mega_) return true;
    if (mega_tried_) return false;
    mega_tried_ = true;
    MegaModelView mv;
    mv.hidden = cfg_.hidden_size; mv.nh = nh_l_; mv.nkv = nkv_l_; mv.hd = cfg_.head_dim; mv.inter = inter_l_;
    mv.max_seq = cfg_.max_seq_len; mv.n_layers = cfg_.n_layers;
    mv.eps = cfg_.norm_eps; mv.theta = cfg_.rope_theta; mv.freq_scale = cfg_.rope_freq_scale;
    mv.tp_rank = tp_rank_; mv.tp_size = tp_size_;
    auto W = [](const Weight& w) { return MegaWeight{w.ptr, w.dtype, w.rows, w.cols, w.pitch}; };
    const size_t kv_stride = (size_t)cfg_.max_seq_len * nkv_l_ * cfg_.head_dim;
    for (int i = 0; i < cfg_.n_layers; i++) {
        const LayerWeights& L = layers_[(size_t)i];
        MegaLayerView lv{L.attn_norm, L.ffn_norm, W(L.wq), W(L.wk), W(L.wv), W(L.wo), W(L.gate), W(L.up), W(L.down),
                         static_cast<uint16_t*>(kc_) + (size_t)i * kv_stride, static_cast<uint16_t*>(vc_) + (size_t)i * kv_stride};
        mv.layers.push_back(lv);
    }
    mv.head = W(head_);
    mv.out_norm = out_norm_;
    mv.logits = tp_size_ == 1 ? logits_ : logits_l_;
    mv.step = step_dev_;
    auto m = std::make_unique<DecodeMega>();
    if (const char* sf = getenv("NT_B200_MEGA_SPLIT_COMPAT"))      // the graph path's split rule: bit-comparable attention
        if (atoi(sf) != 0) m->set_split_fixed(attention_decode_dyn_splits(cfg_.max_seq_len, nh_l_, nkv_l_));
    if (!m->build(mv)) {
        fprintf(stderr, "Note: persistent decode kernel not used (%s); keeping the graph of fused launches\n", m->why().c_str());
        return false;
    }
    if (tp_size_ > 1) {                                            // map every peer's slot/flag allocation (CUDA IPC over NVLink)
        NT_CHECK(comm_ != nullptr, "tensor-parallel model without a communicator");
        const size_t per = DecodeMega::kIpcBytes / sizeof(float);
        float* buf = dmalloc<float>(per * (size_t)(tp_size_ + 1));
        char mine[DecodeMega::kIpcBytes];
        m->export_ipc(mine);
        NT_CUDA_CHECK(cudaMemcpyAsync(buf, mine, sizeof(mine), cudaMemcpyHostToDevice, stream_));
        comm_->all_gather(buf, buf + per, per, stream_);
        std::vector<char> all((size_t)DecodeMega::kIpcBytes * tp_size_);
        NT_CUDA_CHECK(cudaMemcpyAsync(all.data(), buf + per, all.size(), cudaMemcpyDeviceToHost, stream_));
        NT_CUDA_CHECK(cudaStreamSynchronize(stream_));
        NT_CUDA_CHECK(cudaFree(buf));
        m->import_peers(all.data());
    }
    mega_ = std::move(m);
    return true;
}

bool Model::megakernel_active() { return mega_ != nullptr; }

const float* Model::mega_debug_buffer(const char* name, size_t* count) {
    if (!mega_) { if (count) *count = 0; return nullptr; }
    return mega_->debug_buffer(name, count);
}

void Model::mega_trace(bool on) { if (mega_) mega_->set_trace(on); }
size_t Model::mega_trace_read(unsigned long long* out_host, size_t cap) {
    if (!mega_) return 0;
    NT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    return mega_->read_trace(out_host, cap);
}

int Model::mega_plan_kinds(int* kinds, int cap) {
    if (!mega_) return 0;
    const auto& pl = mega_->plan();
    for (int i = 0; i < (int)pl.size() && i < cap; i++) kinds[i] = pl[(size_t)i].kind;
    return (int)pl.size();
}

// embedding gather -> persistent kernel (all layers [+ final norm and LM head]) [-> all-gather of the logits shards]
void Model::run_step_mega(bool with_head) {
    embed_rows(mega_->embed_out(), embd_.ptr, embd_.dtype, step_dev_, 1, cfg_.hidden_size, stream_);
    mega_->launch(with_head, stream_);
    if (with_head && tp_size_ > 1) comm_->all_gather(logits_l_, logits_, (size_t)vocab_l_, stream_);
}

void Model::run_step(bool with_head) {
    // NT_B200_MEGAKERNEL: unset -> the model's own setting (set_use_megakernel, off by default), "0" -> off, anything else -> on
    const char* mk = getenv("NT_B200_MEGAKERNEL");
    const bool want_mega = mk ? env_on("NT_B200_MEGAKERNEL") : use_mega_;
    if (want_mega && ensure_mega()) { run_step_mega(with_head); return; }
    if (!use_graph_ || getenv("NT_B200_NO_GRAPH")) {
        step_body(stream_);
        if (with_head) step_head(stream_);
        return;
    }
    cudaGraphExec_t& g = with_head ? g_full_ : g_body_;
    int& n_kernels = with_head ? n_full_ : n_body_;
    if (!g) {
        cudaGraph_t graph = nullptr;
        const unsigned long long before = launch_count();
        NT_CUDA_CHECK(cudaStreamBeginCapture(stream_, tp_size_ > 1 ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal));
        set_pdl(use_pdl_ && !getenv("NT_B200_NO_PDL"));   // programmatic edges between the step's kernels
        step_body(stream_);
        if (with_head) step_head(stream_);
        set_pdl(false);
        NT_CUDA_CHECK(cudaStreamEndCapture(stream_, &graph));
        NT_CUDA_CHECK(cudaGraphInstantiate(&g, graph, 0));
        NT_CUDA_CHECK(cudaGraphDestroy(graph));
        n_kernels = (int)(launch_count() - before);   // counted once during capture == the first replay below
        NT_CUDA_CHECK(cudaGraphLaunch(g, stream_));
        return;
    }
    NT_CUDA_CHECK(cudaGraphLaunch(g, stream_));
    count_launch(n_kernels);                           // a replay re-runs every captured kernel
}

bool Model::batched_prefill_ok(int seq_len, int start_pos) const {
    if (tp_size_ != 1 || prefill_min_tokens_ <= 0 || seq_len < prefill_min_tokens_ || getenv("NT_B200_NO_BATCHED_PREFILL")) return false;
    if (!attention_prefill_mma_supported(seq_len, nh_l_, nkv_l_, cfg_.head_dim)) return false;
    for (const LayerWeights& L : layers_)
        for (const Weight* w : {&L.wq, &L.wk, &L.wv, &L.wo, &L.gate, &L.up, &L.down})
            if (!dequant_split_supported(w->dtype) || w->rows % 128 != 0 || w->cols % 64 != 0) return false;
    return true;
}

bool Model::f16_direct(const Weight& w) const {
    return w.dtype == DType::F16 && gemm_f16_tc_supported(w.ptr, w.rows, w.cols, w.pitch);
}

void Model::prefill_gemm(float* C, const void* ws, const Weight& w, int T, bool add, cudaStream_t s) {
    if (f16_direct(w)) {
        NT_CHECK(gemm_f16_tc_ws(C, ws, w.ptr, T, w.rows, w.cols, add, s), "prefill GEMM rejected");
        return;
    }
    // quantised (or unaligned F16 / F32) weights: expand once per prompt chunk, then C (+)= A.W_hi^T ; C += A.W_lo^T
    dequant_split(pf_.whi, pf_.wlo, w.ptr, w.dtype, w.pitch, w.rows, w.cols, s);
    NT_CHECK(gemm_f16_tc_ws(C, ws, pf_.whi, T, w.rows, w.cols, add, s), "prefill GEMM (hi) rejected");
    NT_CHECK(gemm_f16_tc_ws(C, ws, pf_.wlo, T, w.rows, w.cols, true, s), "prefill GEMM (lo) rejected");
}

void Model::ensure_prefill_buffers(int tokens) {
    if (tokens <= pf_.cap) return;
    NT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    for (void* p : {(void*)pf_.x, (void*)pf_.q, (void*)pf_.k, (void*)pf_.v, (void*)pf_.attn, pf_.ws, pf_.ws2, (void*)pf_.tok, (void*)pf_.pos,
                    (void*)pf_.g, (void*)pf_.u, pf_.whi, pf_.wlo})
        if (p) cudaFree(p);
    pf_.g = pf_.u = nullptr; pf_.whi = pf_.wlo = nullptr;
    const size_t T = (size_t)((tokens + 127) / 128 * 128);
    const size_t hidden = (size_t)cfg_.hidden_size, qdim = (size_t)nh_l_ * cfg_.head_dim, kvdim = (size_t)nkv_l_ * cfg_.head_dim;
    pf_.x = dmalloc<float>(T * hidden);
    pf_.q = dmalloc<float>(T * qdim);    pf_.k = dmalloc<float>(T * kvdim);   pf_.v = dmalloc<float>(T * kvdim);
    pf_.attn = dmalloc<float>(T * qdim);
    pf_.ws = dmalloc<uint8_t>(gemm_f16_tc_workspace_bytes((int)T, (int)std::max(hidden, qdim)));
    pf_.ws2 = dmalloc<uint8_t>(gemm_f16_tc_workspace_bytes((int)T, inter_l_));
    pf_.tok = dmalloc<int>(T);           pf_.pos = dmalloc<int>(T);
    size_t w_elems = 0;                  // largest matrix that needs the dequantised scratch pair
    bool swiglu_unfused = false;
    for (const LayerWeights& L : layers_) {
        for (const Weight* w : {&L.wq, &L.wk, &L.wv, &L.wo, &L.gate, &L.up, &L.down})
            if (!f16_direct(*w)) w_elems = std::max(w_elems, (size_t)w->rows * (size_t)w->cols);
        swiglu_unfused = swiglu_unfused || !f16_direct(L.gate) || !f16_direct(L.up);
    }
    if (w_elems) { pf_.whi = dmalloc<uint16_t>(w_elems); pf_.wlo = dmalloc<uint16_t>(w_elems); }
    if (swiglu_unfused) { pf_.g = dmalloc<float>(T * (size_t)inter_l_); pf_.u = dmalloc<float>(T * (size_t)inter_l_); }
    pf_.cap = (int)T;
}

void Model::prefill_batched(const int* tokens, int seq_len, int start_pos) {
    cudaStream_t s = stream_;
    const int hidden = cfg_.hidden_size, hd = cfg_.head_dim, max_seq = cfg_.max_seq_len, inter = inter_l_;
    const int qdim = nh_l_ * hd, kvdim = nkv_l_ * hd;
    const float scale = 1.0f / sqrtf((float)hd);
    const size_t kv_stride = (size_t)max_seq * nkv_l_ * hd;
    ensure_prefill_buffers(std::min(seq_len, PREFILL_CHUNK));
    int last_rows = 0;
    for (int c0 = 0, T = 0; c0 < seq_len; c0 += T) {
        // A ragged tail of 1..15 tokens would miss the tiled attention kernel's 16-token minimum (and the per-query kernel it
        // falls back to needs shared memory proportional to the whole context): shorten this chunk by 16 so the tail has >= 17.
        T = std::min(PREFILL_CHUNK, seq_len - c0);
        const int tail = seq_len - c0 - T;
        if (tail > 0 && tail < 16) T -= 16;
        const int p0 = start_pos + c0;
        last_rows = T;
        NT_CUDA_CHECK(cudaMemcpyAsync(pf_.tok, tokens + c0, sizeof(int) * (size_t)T, cudaMemcpyHostToDevice, s));
        iota_kernel<<<(T + 255) / 256, 256, 0, s>>>(pf_.pos, p0, T);
        count_launch();
        embed_rows(pf_.x, embd_.ptr, embd_.dtype, pf_.tok, T, hidden, s);
        for (int i = 0; i < cfg_.n_layers; i++) {
            const LayerWeights& L = layers_[(size_t)i];
            uint16_t* kc = static_cast<uint16_t*>(kc_) + (size_t)i * kv_stride;
            uint16_t* vc = static_cast<uint16_t*>(vc_) + (size_t)i * kv_stride;
            // --- attention sub-block (attention.cpp:120-211 for all T tokens at once) ---
            rmsnorm_split(pf_.ws, pf_.x, L.attn_norm, T, hidden, cfg_.norm_eps, s);
            prefill_gemm(pf_.q, pf_.ws, L.wq, T, false, s);
            prefill_gemm(pf_.k, pf_.ws, L.wk, T, false, s);
            prefill_gemm(pf_.v, pf_.ws, L.wv, T, false, s);
            rope(pf_.q, pf_.k, pf_.pos, T, nh_l_, nkv_l_, hd, cfg_.rope_theta, cfg_.rope_freq_scale, false, s);
            copy_to_kv_cache(kc, vc, pf_.k, pf_.v, T, nkv_l_, hd, p0, max_seq, s);
            attention_prefill(pf_.attn, pf_.q, kc, vc, T, p0, nh_l_, nkv_l_, hd, max_seq, scale, s);
            split_activations(pf_.ws, pf_.attn, T, qdim, s);
            prefill_gemm(pf_.x, pf_.ws, L.wo, T, true, s);
            // --- FFN sub-block (ffn.cpp:85-134) ---
            rmsnorm_split(pf_.ws, pf_.x, L.ffn_norm, T, hidden, cfg_.norm_eps, s);
            if (f16_direct(L.gate) && f16_direct(L.up)) {
                NT_CHECK(gemm_f16_tc_swiglu_ws(pf_.ws2, pf_.ws, L.gate.ptr, L.up.ptr, T, inter, hidden, s), "prefill GEMM (gate/up) rejected");
            } else {                                   // quantised gate/up: plain GEMMs, then SwiGLU and split as separate passes
                prefill_gemm(pf_.g, pf_.ws, L.gate, T, false, s);
                prefill_gemm(pf_.u, pf_.ws, L.up, T, false, s);
                silu_mul(pf_.g, pf_.g, pf_.u, T * inter, s);
                split_activations(pf_.ws2, pf_.g, T, inter, s);
            }
            prefill_gemm(pf_.x, pf_.ws2, L.down, T, true, s);
        }
    }
    copy(hidden_, pf_.x + (size_t)(last_rows - 1) * hidden, hidden, s);       // last token's residual stream -> LM head
    step_head(s);
}

void Model::forward_async(const int* tokens, int seq_len, int start_pos) {
    NT_CHECK(finalized_, "Model::forward before finalize/load");
    NT_CHECK(seq_len >= 1, "forward: empty token list");
    NT_CHECK(start_pos >= 0 && start_pos + seq_len <= cfg_.max_seq_len, "forward: position beyond the KV cache (max_seq_len)");
    for (int t = 0; t < seq_len; t++) NT_CHECK(tokens[t] >= 0 && tokens[t] < cfg_.vocab_size, "token id out of range");
    if (batched_prefill_ok(seq_len, start_pos)) { prefill_batched(tokens, seq_len, start_pos); return; }
    for (int t = 0; t < seq_len; t++) {
        int tok = tokens[t];
        NT_CHECK(tok >= 0 && tok < cfg_.vocab_size, "token id out of range");
        set_step_kernel<<<1, 1, 0, stream_>>>(step_dev_, tok, start_pos + t);
        count_launch();
        run_step(t == seq_len - 1);
    }
}

int Model::sync() {
    if (mega_) mega_->enqueue_abort_read(stream_);
    const cudaError_t e = cudaStreamSynchronize(stream_);
    if (e == cudaSuccess && mega_) mega_->check_abort();
    return (int)e;
}

float* Model::forward(const int* tokens, int seq_len, int start_pos) {
    forward_async(tokens, seq_len, start_pos);
    NT_CUDA_CHECK((cudaError_t)sync());                                      // transformer.cpp:667
    return logits_;
}

int Model::argmax_last() {
    argmax_kernel<<<1, 1024, 0, stream_>>>(logits_, cfg_.vocab_size, argmax_dev_);
    NT_CUDA_CHECK(cudaMemcpyAsync(argmax_host_, argmax_dev_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    NT_CUDA_CHECK((cudaError_t)sync());
    return *argmax_host_;
}

int Model::sample_last(float temperature, int top_k, float top_p, float repeat_penalty, const int* recent_window, int n_window, float r) {
    if (!sample_topk_supported(cfg_.vocab_size, temperature, top_k)) return -1;
    if (n_window > recent_cap_) {
        if (recent_dev_) { NT_CUDA_CHECK(cudaStreamSynchronize(stream_)); NT_CUDA_CHECK(cudaFree(recent_dev_)); }
        recent_cap_ = std::max(256, n_window);
        recent_dev_ = dmalloc<int>((size_t)recent_cap_);
    }
    if (n_window > 0) NT_CUDA_CHECK(cudaMemcpyAsync(recent_dev_, recent_window, sizeof(int) * (size_t)n_window, cudaMemcpyHostToDevice, stream_));
    NT_CHECK(sample_topk(logits_, cfg_.vocab_size, temperature, top_k, top_p, repeat_penalty, recent_dev_, n_window, r, argmax_dev_, stream_),
             "sample_topk rejected supported settings");
    NT_CUDA_CHECK(cudaMemcpyAsync(argmax_host_, argmax_dev_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    NT_CUDA_CHECK((cudaError_t)sync());
    return *argmax_host_;
}

}}  // namespace nt::b200
