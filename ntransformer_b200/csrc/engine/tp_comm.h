// tp_comm.h — tensor-parallel collectives over NCCL (NVLink 5 / NVSwitch).
// New work relative to the reference, which is single-GPU.  One process per GPU: the
// 128-byte ncclUniqueId is created on rank 0 and handed to the other ranks by the host plumbing
// (torch.distributed in bench.py); libnccl.so.2 is resolved at run time (it is already in the process when
// torch is imported) so libnt_b200.so has no link-time NCCL dependency on single-GPU boxes.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace nt { namespace b200 {

class TPComm {
public:
    static constexpr int kIdBytes = 128;
    static bool unique_id(void* out128);                       // rank 0 only
    TPComm() = default;
    ~TPComm();
    bool init(const void* id128, int rank, int size);          // collective: every rank must call it
    void all_reduce_sum(float* buf, size_t n, cudaStream_t s); // in place
    void all_gather(const float* send, float* recv, size_t n_per_rank, cudaStream_t s);
    int rank() const { return rank_; }
    int size() const { return size_; }
private:
    void* comm_ = nullptr;
    int rank_ = 0, size_ = 1;
};

}}  // namespace nt::b200
