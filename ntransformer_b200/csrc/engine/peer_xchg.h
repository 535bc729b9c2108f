// peer_xchg.h — tensor-parallel sum of the o-projection / down-projection partials over NVLink peer memory, without NCCL.
//
// New relative to the single-GPU reference ("one all-reduce per sub-block").  One-shot all-reduce in two kernels:
//   * the producing GEMV (gemv_kquant.cu, epilogue GEMV_PEER) stores each finished row of its partial result straight into the
//     "slot" [parity][source rank][hidden] of EVERY rank (stores to cudaIpc-mapped peer memory, i.e. NVLink writes issued from
//     the GEMV's epilogue while other CTAs are still streaming weights).  A slot element is one 64-bit word {sequence : value}
//     written by a single store, so the value is its own arrival flag (the scheme of NCCL's LL protocol): no fence and no
//     separate flag hop;
//   * xchg_reduce_kernel (this file): the thread that owns element e polls its tp slot elements until they carry the sequence
//     number, then hidden[e] += sum_r slot[r][e] in rank order on every rank (deterministic, bit-identical everywhere: the
//     replicas' greedy ids stay in lock-step).  A first version with plain rows + system fence + flag line per rank measured
//     10.87 ms/token at TP-2 against 9.38 with ncclAllReduce (profiles/r02_tp2_*): the fence and flag hop cost more than NCCL.
// Slots are double-buffered by the sequence number's parity: a rank can run at most one exchange ahead of a peer.
// Every spin has a time-out that raises an abort word instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <vector>
#include "../kernels_internal.h"

namespace nt { namespace b200 {

class TPComm;

class PeerXchg {
public:
    PeerXchg() = default;
    ~PeerXchg();
    PeerXchg(const PeerXchg&) = delete;
    PeerXchg& operator=(const PeerXchg&) = delete;
    // Collective: allocates this rank's buffer, exchanges the IPC handles through `comm` (one NCCL all-gather at set-up) and maps
    // the peers.  Returns false (with a note on stderr) when peer memory cannot be mapped: the caller keeps NCCL.
    bool init(TPComm* comm, int rank, int size, int hidden, cudaStream_t s);
    const PeerOut& out() const { return out_; }          // what the GEMV epilogue needs
    // hidden += sum over ranks of the partials of the exchange in flight; ends the exchange.
    void reduce_residual(float* hidden, cudaStream_t s);
    void enqueue_abort_read(cudaStream_t s);
    bool aborted() const;                                 // valid after the stream has been synchronised
private:
    PeerOut out_{};
    void* local_ = nullptr;
    std::vector<void*> peer_maps_;
    unsigned* reduce_arrive_ = nullptr;
    unsigned* abort_host_ = nullptr;
    unsigned long long timeout_ns_ = 2000000000ull;
};

}}  // namespace nt::b200
