// model.cu — see model.h.
#include "model.h"
#include "tp_comm.h"
#include "peer_xchg.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <thread>
#include <unistd.h>

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
                    (void*)attn_tickets_,
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

// ---- GGUF -> HBM ----------------------------------------------------------------------------------------------------------------
// The reference copies one tensor at a time with a synchronous cudaMemcpy straight from the pageable mmap
// (transformer.cpp:286-328 via tensor.cpp:224-225): the driver stages every page through its own bounce buffer on one thread.
// Here: every tensor's shard is planned first (shapes validated against the config — a mismatching file is rejected, not
// trusted), all weights live in ONE device allocation, and NT_LOAD_THREADS reader threads each cycle
//     pread(chunk of whole rows) -> pinned staging buffer -> cudaMemcpyAsync / cudaMemcpy2DAsync on the thread's own stream
// so page-cache reads, PCIe copies and the other threads' reads overlap.  Under tensor parallelism a rank uploads only its
// shard: row shards are one contiguous file range, column shards (attn_output, ffn_down) read whole rows and copy the rank's
// byte slice of each row (block-aligned, padded to a 16-byte pitch for TMA).
namespace {

struct UploadJob {
    const GGUFTensorInfo* ti = nullptr;
    size_t file_off = 0;        // first byte of the first row this rank needs
    size_t src_pitch = 0;       // bytes between rows in the file
    size_t col_off = 0;         // byte offset of the rank's slice inside a row
    size_t width = 0;           // bytes of a row this rank keeps
    size_t rows = 0;
    size_t dst_pitch = 0;
    size_t dst_off = 0;         // inside the arena
};

struct Chunk { int job; size_t row0, nrows; };

}  // namespace

bool Model::load_gguf(const std::string& path, int max_context, int tp_rank, int tp_size) {
    fprintf(stderr, "Loading model: %s\n", path.c_str());
    const auto t_start = std::chrono::steady_clock::now();
    GGUFFile f;
    if (!f.open(path)) return false;
    ModelConfig cfg = f.config();
    if (cfg.max_seq_len > max_context) {          // transformer.cpp:70-74
        fprintf(stderr, "Note: Capping context from %d to %d tokens (use --ctx-size to change)\n", cfg.max_seq_len, max_context);
        cfg.max_seq_len = max_context;
    }
    cfg.print();
    f.print_info();
    if (cfg.hidden_size <= 0 || cfg.n_layers <= 0 || cfg.n_heads <= 0 || cfg.n_kv_heads <= 0 || cfg.vocab_size <= 0 ||
        cfg.intermediate_size <= 0 || cfg.n_heads % tp_size || cfg.n_kv_heads % tp_size || cfg.intermediate_size % tp_size) {
        fprintf(stderr, "GGUF: model dimensions are invalid (or do not divide by the tensor-parallel size %d)\n", tp_size);
        return false;
    }
    vocab_ = f.vocab();
    init(cfg, tp_rank, tp_size);

    // ---- plan: one job per tensor, shapes checked against the config ----
    const int hidden = cfg_.hidden_size, hd = cfg_.head_dim;
    std::vector<UploadJob> jobs;
    std::vector<Weight*> job_w;
    std::vector<const float**> job_f;
    size_t arena = 0;
    bool ok = true;
    auto plan = [&](const std::string& name, int rows, int cols, int split, Weight* w, const float** fptr) {
        const GGUFTensorInfo* ti = f.find(name);
        if (!ti) { fprintf(stderr, "GGUF: tensor not found: %s\n", name.c_str()); ok = false; return; }
        const long long frows = ti->shape.size() > 1 ? ti->shape[1] : 1, fcols = ti->shape[0];
        long long extra = 1;
        for (size_t d = 2; d < ti->shape.size(); d++) extra *= ti->shape[d];
        if (fcols != cols || frows != rows || extra != 1) {
            fprintf(stderr, "GGUF: tensor %s has shape [%lld x %lld], the config needs [%d x %d]\n", name.c_str(), frows, fcols, rows, cols);
            ok = false; return;
        }
        if (fptr && ti->dtype != DType::F32) { fprintf(stderr, "GGUF: %s must be F32\n", name.c_str()); ok = false; return; }
        UploadJob j;
        j.ti = ti;
        const DType dt = ti->dtype;
        const size_t row_bytes = dtype_row_size(dt, (size_t)cols);
        j.src_pitch = row_bytes;
        size_t r0 = 0, nrows = (size_t)rows;
        j.width = row_bytes; j.dst_pitch = row_bytes;
        int w_rows = rows, w_cols = cols;
        if (tp_size_ > 1 && split == 1) {
            const int per = (rows + tp_size_ - 1) / tp_size_;
            r0 = (size_t)std::min(rows, tp_rank_ * per);
            nrows = (size_t)std::min(rows, (int)r0 + per) - r0;
            w_rows = (int)nrows;
        } else if (tp_size_ > 1 && split == 2) {
            const int bs = (int)dtype_block_size(dt);
            if (cols % tp_size_ != 0 || (cols / tp_size_) % bs != 0) {
                fprintf(stderr, "GGUF: %s: a %d-way column shard is not quantisation-block aligned\n", name.c_str(), tp_size_);
                ok = false; return;
            }
            w_cols = cols / tp_size_;
            j.width = dtype_row_size(dt, (size_t)w_cols);
            j.col_off = (size_t)tp_rank_ * j.width;
            j.dst_pitch = round16(j.width);                     // rows start 16 B aligned for the TMA path
        }
        j.rows = nrows;
        j.file_off = f.data_offset() + (size_t)ti->offset + r0 * row_bytes;
        j.dst_off = arena;
        arena += (std::max<size_t>(j.dst_pitch * nrows, 16) + 255) & ~(size_t)255;
        if (w) { w->dtype = dt; w->rows = w_rows; w->cols = w_cols; w->pitch = j.dst_pitch; w->owned = true; }
        jobs.push_back(j); job_w.push_back(w); job_f.push_back(fptr);
    };
    plan("token_embd.weight", cfg_.vocab_size, hidden, 0, &embd_, nullptr);
    const bool has_head = f.find("output.weight") != nullptr;
    if (has_head) plan("output.weight", cfg_.vocab_size, hidden, 1, &head_, nullptr);
    else if (tp_size_ > 1) plan("token_embd.weight", cfg_.vocab_size, hidden, 1, &head_, nullptr);   // tied embeddings, sharded head
    plan("output_norm.weight", 1, hidden, 0, nullptr, &out_norm_);
    for (int i = 0; i < cfg_.n_layers && ok; i++) {
        const std::string p = "blk." + std::to_string(i) + ".";
        LayerWeights& L = layers_[(size_t)i];
        plan(p + "attn_norm.weight", 1, hidden, 0, nullptr, &L.attn_norm);
        plan(p + "attn_q.weight", cfg_.n_heads * hd, hidden, 1, &L.wq, nullptr);
        plan(p + "attn_k.weight", cfg_.n_kv_heads * hd, hidden, 1, &L.wk, nullptr);
        plan(p + "attn_v.weight", cfg_.n_kv_heads * hd, hidden, 1, &L.wv, nullptr);
        plan(p + "attn_output.weight", hidden, cfg_.n_heads * hd, 2, &L.wo, nullptr);
        plan(p + "ffn_norm.weight", 1, hidden, 0, nullptr, &L.ffn_norm);
        plan(p + "ffn_gate.weight", cfg_.intermediate_size, hidden, 1, &L.gate, nullptr);
        plan(p + "ffn_up.weight", cfg_.intermediate_size, hidden, 1, &L.up, nullptr);
        plan(p + "ffn_down.weight", hidden, cfg_.intermediate_size, 2, &L.down, nullptr);
    }
    if (!ok) return false;
    if (embd_.dtype == DType::Q5_K)
        fprintf(stderr, "Error: Unsupported embedding dtype: Q5_K (rows read as zeros, like the reference)\n");

    uint8_t* base = nullptr;
    if (cudaMalloc(&base, std::max<size_t>(arena, 256)) != cudaSuccess) {
        cudaGetLastError();
        fprintf(stderr, "model: cannot allocate %.2f GB of device memory for the weights\n", arena / 1e9);
        return false;
    }
    owned_.push_back(base);
    size_t padded = 0;
    for (size_t k = 0; k < jobs.size(); k++) {
        if (job_w[k]) job_w[k]->ptr = base + jobs[k].dst_off;
        if (job_f[k]) *job_f[k] = reinterpret_cast<const float*>(base + jobs[k].dst_off);
        if (jobs[k].dst_pitch != jobs[k].width) padded += jobs[k].dst_pitch * jobs[k].rows;
    }
    if (padded) NT_CUDA_CHECK(cudaMemset(base, 0, arena));          // pitch padding of column shards must be finite bytes
    if (!has_head && tp_size_ == 1) { head_ = embd_; head_.owned = false; }   // tied embeddings (transformer.cpp:96-99)

    // ---- chunks of whole rows, at most CHUNK bytes of file each ----
    constexpr size_t CHUNK = (size_t)32 << 20;
    std::vector<Chunk> chunks;
    size_t file_bytes = 0;
    for (size_t k = 0; k < jobs.size(); k++) {
        const UploadJob& j = jobs[k];
        if (j.src_pitch > CHUNK) { fprintf(stderr, "model: a tensor row exceeds the staging chunk\n"); return false; }
        const size_t per = std::max<size_t>(1, CHUNK / j.src_pitch);
        for (size_t r = 0; r < j.rows; r += per) chunks.push_back({(int)k, r, std::min(per, j.rows - r)});
        file_bytes += j.rows * j.src_pitch;
    }
    int n_threads = 8;
    if (const char* e = getenv("NT_LOAD_THREADS")) n_threads = std::max(1, std::min(32, atoi(e)));
    n_threads = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, chunks.size()));
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    const int fd = f.fd();
    int dev = 0;
    NT_CUDA_CHECK(cudaGetDevice(&dev));
    auto reader = [&]() {
        cudaSetDevice(dev);
        uint8_t* stage[2] = {nullptr, nullptr};
        cudaEvent_t done[2];
        cudaStream_t st = nullptr;
        bool up = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        for (int b = 0; b < 2 && up; b++)
            up = cudaHostAlloc((void**)&stage[b], CHUNK, cudaHostAllocDefault) == cudaSuccess && cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) == cudaSuccess;
        if (!up) { failed = 1; return; }
        int b = 0;
        for (size_t c = next.fetch_add(1); c < chunks.size() && !failed; c = next.fetch_add(1)) {
            const Chunk& ch = chunks[c];
            const UploadJob& j = jobs[(size_t)ch.job];
            if (cudaEventSynchronize(done[b]) != cudaSuccess) { failed = 1; break; }      // the copy that last used this buffer
            const size_t bytes = ch.nrows * j.src_pitch;
            size_t got = 0;
            while (got < bytes) {
                const ssize_t n = pread(fd, stage[b] + got, bytes - got, (off_t)(j.file_off + ch.row0 * j.src_pitch + got));
                if (n <= 0) { failed = 2; break; }
                got += (size_t)n;
            }
            if (failed) break;
            uint8_t* dst = base + j.dst_off + ch.row0 * j.dst_pitch;
            cudaError_t e;
            if (j.width == j.src_pitch && j.dst_pitch == j.src_pitch) e = cudaMemcpyAsync(dst, stage[b], bytes, cudaMemcpyHostToDevice, st);
            else e = cudaMemcpy2DAsync(dst, j.dst_pitch, stage[b] + j.col_off, j.src_pitch, j.width, ch.nrows, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess || cudaEventRecord(done[b], st) != cudaSuccess) { failed = 1; break; }
            b ^= 1;
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) failed = 1;
        for (int k = 0; k < 2; k++) { if (stage[k]) cudaFreeHost(stage[k]); cudaEventDestroy(done[k]); }
        cudaStreamDestroy(st);
    };
    const auto t_copy = std::chrono::steady_clock::now();
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; t++) pool.emplace_back(reader);
        for (std::thread& t : pool) t.join();
    }
    if (failed) {
        cudaGetLastError();
        fprintf(stderr, failed == 2 ? "model: short read from %s\n" : "model: staging / copy failure while loading %s\n", path.c_str());
        return false;
    }
    const double copy_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_copy).count();
    fprintf(stderr, "  Loaded %d layers: %.2f GB read, %.2f GB resident on this GPU, %.2f s (%.1f GB/s, %d reader threads)\n", cfg_.n_layers,
            file_bytes / 1e9, arena / 1e9, copy_s, file_bytes / 1e9 / std::max(copy_s, 1e-9), n_threads);
    if (!finalize()) return false;
    load_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
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
    attn_tickets_ = dmalloc<unsigned>((size_t)std::max(1, nh_l_));
    NT_CUDA_CHECK(cudaMemset(attn_tickets_, 0, (size_t)std::max(1, nh_l_) * sizeof(unsigned)));
    if (const char* fz = getenv("NT_B200_FUSE")) fuse_mask_ = atoi(fz);
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
    // (the unfused path; with the peer exchange mapped the short chain sums through NVLink stores instead, see step_body)
    comm_->all_reduce_sum(partial, (size_t)cfg_.hidden_size, s);
    add_inplace(hidden_, partial, cfg_.hidden_size, s);
}

// The short launch chain needs every projection on the TMA/dp4a GEMV (its prologue carries the norm + quantiser).
// Collective (every rank takes the same branch): map the peers' exchange buffers once, before any graph is captured.
void Model::ensure_xchg() {
    if (tp_size_ == 1 || xchg_tried_) return;
    xchg_tried_ = true;
    NT_CHECK(comm_ != nullptr, "tensor-parallel model without a communicator");
    if (env_on("NT_B200_TP_NCCL")) return;
    auto x = std::make_unique<PeerXchg>();
    float ok = x->init(comm_, tp_rank_, tp_size_, cfg_.hidden_size, stream_) ? 0.f : 1.f;
    float* flag = dmalloc<float>(1);                       // all ranks agree: any failure anywhere keeps NCCL everywhere
    NT_CUDA_CHECK(cudaMemcpyAsync(flag, &ok, sizeof(float), cudaMemcpyHostToDevice, stream_));
    comm_->all_reduce_sum(flag, 1, stream_);
    NT_CUDA_CHECK(cudaMemcpyAsync(&ok, flag, sizeof(float), cudaMemcpyDeviceToHost, stream_));
    NT_CUDA_CHECK(cudaStreamSynchronize(stream_));
    NT_CUDA_CHECK(cudaFree(flag));
    if (ok == 0.f) xchg_ = std::move(x);
    else if (tp_rank_ == 0) fprintf(stderr, "Note: NVLink peer exchange unavailable; tensor-parallel sums go through ncclAllReduce\n");
}

bool Model::chain_ok() const {
    if ((tp_size_ != 1 && !xchg_) || !(fuse_mask_ & 1)) return false;
    if (cfg_.hidden_size % 128 != 0 || inter_l_ % 128 != 0 || (nh_l_ * cfg_.head_dim) % 128 != 0) return false;
    auto kq = [&](const Weight& w, int K) {
        GemvMat m; m.W = w.ptr; m.y = nullptr; m.out = w.rows; m.dtype = w.dtype; m.row_pitch = w.pitch;
        return w.cols == K && gemv_kq_supported(&m, 1, K);
    };
    for (const LayerWeights& L : layers_) {
        GemvMat qkv[3];
        const Weight* ws[3] = {&L.wq, &L.wk, &L.wv};
        for (int i = 0; i < 3; i++) { qkv[i].W = ws[i]->ptr; qkv[i].out = ws[i]->rows; qkv[i].dtype = ws[i]->dtype; qkv[i].row_pitch = ws[i]->pitch; }
        if (!gemv_kq_supported(qkv, 3, cfg_.hidden_size)) return false;
        if (!kq(L.wo, nh_l_ * cfg_.head_dim) || !kq(L.gate, cfg_.hidden_size) || !kq(L.up, cfg_.hidden_size) || !kq(L.down, inter_l_)) return false;
        if (L.gate.dtype != L.up.dtype) return false;
    }
    return kq(head_, cfg_.hidden_size) || head_.rows == 0;
}

// One decoded token, all layers.  Two forms of the reference's launch sequence (Attention::forward attention.cpp:120-211,
// FFN::forward ffn.cpp:85-134, 15 launches per layer there):
//   short chain (one GPU, all projections on the TMA/dp4a GEMV) — 6 launches per layer:
//     [RMSNorm + quantise + q/k/v GEMV] [RoPE + KV write + attention + merge] [quantise + o GEMV += residual]
//     [RMSNorm + quantise + gate/up GEMV + SwiGLU] [quantise] [down GEMV += residual]
//     The norms and the quantisers of hidden-sized vectors run in the consuming GEMV's prologue while its first weights are
//     already in flight (gemv_kquant.cu); only the intermediate-sized activation vector keeps its own quantise launch;
//   unfused — 8-10 launches per layer (tensor parallel, mixed dtypes, NT_B200_FUSE=0).
void Model::step_body(cudaStream_t s) {
    const int hidden = cfg_.hidden_size, hd = cfg_.head_dim, max_seq = cfg_.max_seq_len;
    const float scale = 1.0f / sqrtf((float)hd);                         // attention.cpp:21
    const size_t kv_stride = (size_t)max_seq * nkv_l_ * hd;               // transformer.cpp:629
    const int* pos_dev = step_dev_ + 1;
    const bool chain = chain_ok();
    const bool attn1 = (fuse_mask_ & 2) != 0;
    embed_rows(hidden_, embd_.ptr, embd_.dtype, step_dev_, 1, hidden, s);
    auto mat = [](const Weight& w, float* y) { GemvMat m; m.W = w.ptr; m.y = y; m.out = w.rows; m.dtype = w.dtype; m.row_pitch = w.pitch; return m; };
    for (int i = 0; i < cfg_.n_layers; i++) {
        const LayerWeights& L = layers_[(size_t)i];
        uint16_t* kc = static_cast<uint16_t*>(kc_) + (size_t)i * kv_stride;
        uint16_t* vc = static_cast<uint16_t*>(vc_) + (size_t)i * kv_stride;
        // --- attention sub-block ---
        if (chain) {
            GemvMat m3[3] = {mat(L.wq, q_), mat(L.wk, k_), mat(L.wv, v_)};
            GemvInput in; in.x = hidden_; in.norm_w = L.attn_norm; in.eps = cfg_.norm_eps;
            gemv_kq(m3, 3, hidden, in, GEMV_STORE, s);
        } else {
            const Weight* ws[3] = {&L.wq, &L.wk, &L.wv}; float* ys[3] = {q_, k_, v_}; matvec(ws, ys, 3, hidden_, L.attn_norm, GEMV_STORE, s);
        }
        const bool o_xq = !chain && o_xq_fusable(L.wo);            // unfused path: attention emits xq for the o-projection
        if (attn1) {
            attention_decode_fused(attn_, q_, k_, v_, kc, vc, pos_dev, max_seq, nh_l_, nkv_l_, hd, cfg_.rope_theta, cfg_.rope_freq_scale, scale,
                                   attn_scratch_, attn_tickets_, o_xq ? xq_a_ : nullptr, s);
        } else if (fuse_mask_ & 4) {
            attention_decode_rope_dyn(attn_, q_, k_, v_, kc, vc, pos_dev, max_seq, nh_l_, nkv_l_, hd, cfg_.rope_theta, cfg_.rope_freq_scale, scale,
                                      attn_scratch_, o_xq ? xq_a_ : nullptr, s);
        } else {
            rope_kv_decode(q_, k_, v_, kc, vc, pos_dev, nh_l_, nkv_l_, hd, cfg_.rope_theta, cfg_.rope_freq_scale, max_seq, s);
            attention_decode_dyn(attn_, q_, kc, vc, pos_dev, max_seq, nh_l_, nkv_l_, hd, scale, attn_scratch_, o_xq ? xq_a_ : nullptr, s);
        }
        if (chain) {
            GemvMat m = mat(L.wo, hidden_);
            GemvInput in; in.x = attn_;
            if (tp_size_ == 1) gemv_kq(&m, 1, L.wo.cols, in, GEMV_ADD, s);
            else { gemv_kq_peer(m, L.wo.cols, in, xchg_->out(), s); xchg_->reduce_residual(hidden_, s); }
        } else {
            const Weight* ws[1] = {&L.wo};
            float* ys[1] = {tp_size_ == 1 ? hidden_ : part_};
            const GemvEpilogue ep = tp_size_ == 1 ? GEMV_ADD : GEMV_STORE;
            if (o_xq) {
                GemvMat m = mat(L.wo, ys[0]);
                gemv_kq(&m, 1, L.wo.cols, xq_a_, ep, s);
            } else {
                matvec(ws, ys, 1, attn_, nullptr, ep, s);
            }
            if (tp_size_ > 1) reduce_residual(part_, s);
        }
        // --- FFN sub-block ---
        if (chain) {
            GemvMat m2[2] = {mat(L.gate, act_), mat(L.up, up_)};
            GemvInput in; in.x = hidden_; in.norm_w = L.ffn_norm; in.eps = cfg_.norm_eps;
            gemv_kq(m2, 2, hidden, in, GEMV_SWIGLU, s);
            // The activation vector's quantiser: in the down GEMV's prologue when its staging (3.4 x inter bytes) fits over the ring
            // (tensor-parallel shards: inter / tp <= 8192), as a launch of its own for the full 28672-wide vector.
            GemvMat md = mat(L.down, hidden_);
            GemvInput ind;
            if (inter_l_ <= 8192) ind.x = act_;
            else { quantize_x(act_, xq_i_, inter_l_, s); ind.xq = xq_i_; }
            if (tp_size_ == 1) gemv_kq(&md, 1, inter_l_, ind, GEMV_ADD, s);
            else { gemv_kq_peer(md, inter_l_, ind, xchg_->out(), s); xchg_->reduce_residual(hidden_, s); }
            continue;
        }
        { const Weight* ws[2] = {&L.gate, &L.up}; float* ys[2] = {act_, up_}; matvec(ws, ys, 2, hidden_, L.ffn_norm, GEMV_SWIGLU, s); }
        { const Weight* ws[1] = {&L.down};
          if (tp_size_ == 1) { float* ys[1] = {hidden_}; matvec(ws, ys, 1, act_, nullptr, GEMV_ADD, s); }
          else { float* ys[1] = {part_}; matvec(ws, ys, 1, act_, nullptr, GEMV_STORE, s); reduce_residual(part_, s); } }
    }
}

void Model::step_head(cudaStream_t s, bool from_chain) {
    const Weight* ws[1] = {&head_};                                               // final norm fused: transformer.cpp:657-665
    if (from_chain) {
        GemvMat m; m.W = head_.ptr; m.y = tp_size_ == 1 ? logits_ : logits_l_; m.out = head_.rows; m.dtype = head_.dtype; m.row_pitch = head_.pitch;
        GemvInput in; in.x = hidden_; in.norm_w = out_norm_; in.eps = cfg_.norm_eps;
        if (head_.rows > 0) gemv_kq(&m, 1, cfg_.hidden_size, in, GEMV_STORE, s);
        if (tp_size_ > 1) comm_->all_gather(logits_l_, logits_, (size_t)vocab_l_, s);
        return;
    }
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
    ensure_xchg();
    if (!use_graph_ || getenv("NT_B200_NO_GRAPH")) {
        step_body(stream_);
        if (with_head) step_head(stream_, chain_ok());
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
        if (with_head) step_head(stream_, chain_ok());
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
    if (xchg_) xchg_->enqueue_abort_read(stream_);
    const cudaError_t e = cudaStreamSynchronize(stream_);
    if (e == cudaSuccess && mega_) mega_->check_abort();
    if (e == cudaSuccess && xchg_) NT_CHECK(!xchg_->aborted(), "tensor-parallel peer exchange timed out (a rank never published its partial rows)");
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
