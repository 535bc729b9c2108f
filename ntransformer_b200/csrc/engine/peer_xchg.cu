// peer_xchg.cu — see peer_xchg.h.
#include "peer_xchg.h"
#include "tp_comm.h"
#include "../ring.cuh"

#include <cstdlib>
#include <cstring>

namespace nt { namespace b200 {

namespace {

constexpr int LINE = 32;                                 // words per 128-byte line: one counter per line

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// hidden[e] += slot[parity][0][e] + ... + slot[parity][tp-1][e]  (rank order).  Every slot element is {sequence : value}: the thread
// that owns element e polls its tp elements (all loads in flight together) until each carries sequence number s — the peers'
// GEMV epilogues deliver them with single 64-bit NVLink stores, so there is no fence and no flag round trip on the path.
__global__ void __launch_bounds__(256) xchg_reduce_kernel(PeerOut P, float* __restrict__ hidden, unsigned* __restrict__ reduce_arrive,
                                                          unsigned long long timeout_ns) {
    pdl_launch_dependents();
    pdl_wait();
    const unsigned s = __ldcg(P.seq) + 1u;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < P.hidden) {
        const unsigned long long* sl = P.slots[P.rank] + (size_t)(s & 1u) * P.size * (size_t)P.hidden + e;
        unsigned long long pk[PeerOut::kMaxTP];
        unsigned pending = (1u << P.size) - 1u;
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (pending) {
#pragma unroll
            for (int r = 0; r < PeerOut::kMaxTP; r++)
                if (pending & (1u << r)) pk[r] = ld_relaxed_sys_u64(sl + (size_t)r * P.hidden);
#pragma unroll
            for (int r = 0; r < PeerOut::kMaxTP; r++)
                if ((pending & (1u << r)) && (unsigned)(pk[r] >> 32) == s) pending &= ~(1u << r);
            if (pending && (++spins & 63u) == 0) {
                if (!t0) t0 = global_timer_ns();
                if (__ldcg(P.abort_word) || global_timer_ns() - t0 > timeout_ns) {
                    atomicExch(P.abort_word, 1u + (unsigned)__ffs((int)pending) - 1u);     // which peer never arrived (+1)
                    break;
                }
            }
        }
        float t = __uint_as_float((unsigned)pk[0]);
#pragma unroll
        for (int r = 1; r < PeerOut::kMaxTP; r++) if (r < P.size) t += __uint_as_float((unsigned)pk[r]);
        hidden[e] += t;
    }
    // the last CTA ends the exchange: the next producer sends s + 1 (every CTA has read *seq above)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(reduce_arrive, 1u);
        if (prev == gridDim.x - 1) { *reduce_arrive = 0; *P.seq = s; }
    }
}

}  // namespace

PeerXchg::~PeerXchg() {
    for (void* p : peer_maps_) if (p) cudaIpcCloseMemHandle(p);
    if (local_) cudaFree(local_);
    if (reduce_arrive_) cudaFree(reduce_arrive_);
    if (abort_host_) cudaFreeHost(abort_host_);
}

bool PeerXchg::init(TPComm* comm, int rank, int size, int hidden, cudaStream_t s) {
    if (!comm || size < 2 || size > PeerOut::kMaxTP) return false;
    if (const char* t = getenv("NT_B200_XCHG_TIMEOUT_MS")) timeout_ns_ = (unsigned long long)atoll(t) * 1000000ull;
    const size_t slot_words = (size_t)2 * size * hidden * 2;              // 64-bit elements
    const size_t words = slot_words + (size_t)2 * LINE;                   // slots | seq | abort
    NT_CUDA_CHECK(cudaMalloc(&local_, words * 4));
    NT_CUDA_CHECK(cudaMemset(local_, 0, words * 4));
    NT_CUDA_CHECK(cudaMalloc(&reduce_arrive_, 128));
    NT_CUDA_CHECK(cudaMemset(reduce_arrive_, 0, 128));
    NT_CUDA_CHECK(cudaMallocHost(&abort_host_, sizeof(unsigned)));
    *abort_host_ = 0;
    NT_CUDA_CHECK(cudaDeviceSynchronize());                              // the zeroes are in place before any peer can write
    // exchange the IPC handles (64 bytes each) through the communicator
    constexpr size_t HB = sizeof(cudaIpcMemHandle_t);
    static_assert(HB % 4 == 0, "IPC handle size");
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, local_) != cudaSuccess) { cudaGetLastError(); fprintf(stderr, "PeerXchg: cudaIpcGetMemHandle failed; keeping NCCL\n"); return false; }
    float* buf = nullptr;
    NT_CUDA_CHECK(cudaMalloc(&buf, HB * (size_t)(size + 1)));
    NT_CUDA_CHECK(cudaMemcpyAsync(buf, &mine, HB, cudaMemcpyHostToDevice, s));
    comm->all_gather(buf, buf + HB / 4, HB / 4, s);
    std::vector<char> all(HB * (size_t)size);
    NT_CUDA_CHECK(cudaMemcpyAsync(all.data(), buf + HB / 4, all.size(), cudaMemcpyDeviceToHost, s));
    NT_CUDA_CHECK(cudaStreamSynchronize(s));
    NT_CUDA_CHECK(cudaFree(buf));
    peer_maps_.assign((size_t)size, nullptr);
    out_ = PeerOut{};
    out_.rank = rank; out_.size = size; out_.hidden = hidden;
    for (int r = 0; r < size; r++) {
        void* p = local_;
        if (r != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + (size_t)r * HB, HB);
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                fprintf(stderr, "PeerXchg: cannot map rank %d's buffer (cudaIpcOpenMemHandle); keeping NCCL\n", r);
                return false;
            }
            peer_maps_[(size_t)r] = p;
        }
        out_.slots[r] = static_cast<unsigned long long*>(p);
    }
    unsigned* base = static_cast<unsigned*>(local_) + slot_words;
    out_.seq = base;
    out_.abort_word = base + LINE;
    return true;
}

void PeerXchg::reduce_residual(float* hidden, cudaStream_t s) {
    launch_k(xchg_reduce_kernel, dim3((unsigned)((out_.hidden + 255) / 256)), dim3(256), 0, s, out_, hidden, reduce_arrive_, timeout_ns_);
    count_launch();
}

void PeerXchg::enqueue_abort_read(cudaStream_t s) {
    NT_CUDA_CHECK(cudaMemcpyAsync(abort_host_, out_.abort_word, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
}
bool PeerXchg::aborted() const { return abort_host_ && *abort_host_ != 0; }

}}  // namespace nt::b200
