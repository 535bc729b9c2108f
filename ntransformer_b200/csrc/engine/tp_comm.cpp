#include "tp_comm.h"
#include "../nt_types.h"
#include <dlfcn.h>
#include <cstring>

namespace nt { namespace b200 {
namespace {
struct NcclId { char bytes[128]; };
using fn_get_id = int (*)(NcclId*);
using fn_init = int (*)(void**, int, NcclId, int);
using fn_destroy = int (*)(void*);
using fn_allreduce = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using fn_allgather = int (*)(const void*, void*, size_t, int, void*, cudaStream_t);
using fn_errstr = const char* (*)(int);
struct Api {
    fn_get_id get_id = nullptr; fn_init init = nullptr; fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr; fn_allgather allgather = nullptr; fn_errstr errstr = nullptr;
    bool ok = false;
};
Api& api() {
    static Api a;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { fprintf(stderr, "TPComm: cannot load libnccl.so.2: %s\n", dlerror()); return a; }
        a.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
        a.init = (fn_init)dlsym(h, "ncclCommInitRank");
        a.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
        a.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
        a.allgather = (fn_allgather)dlsym(h, "ncclAllGather");
        a.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
        a.ok = a.get_id && a.init && a.destroy && a.allreduce && a.allgather;
    }
    return a;
}
constexpr int kNcclFloat32 = 7, kNcclSum = 0;
void check(int rc, const char* what) {
    if (rc != 0) {
        fprintf(stderr, "NCCL error in %s: %s\n", what, api().errstr ? api().errstr(rc) : "?");
        abort();
    }
}
}  // namespace

bool TPComm::unique_id(void* out128) {
    if (!api().ok) return false;
    NcclId id;
    if (api().get_id(&id) != 0) return false;
    memcpy(out128, id.bytes, 128);
    return true;
}
bool TPComm::init(const void* id128, int rank, int size) {
    if (!api().ok) return false;
    NcclId id;
    memcpy(id.bytes, id128, 128);
    rank_ = rank; size_ = size;
    int rc = api().init(&comm_, size, id, rank);
    if (rc != 0) { fprintf(stderr, "ncclCommInitRank failed: %s\n", api().errstr ? api().errstr(rc) : "?"); comm_ = nullptr; return false; }
    return true;
}
TPComm::~TPComm() { if (comm_) api().destroy(comm_); }
void TPComm::all_reduce_sum(float* buf, size_t n, cudaStream_t s) {
    check(api().allreduce(buf, buf, n, kNcclFloat32, kNcclSum, comm_, s), "ncclAllReduce");
}
void TPComm::all_gather(const float* send, float* recv, size_t n, cudaStream_t s) {
    check(api().allgather(send, recv, n, kNcclFloat32, comm_, s), "ncclAllGather");
}
}}  // namespace nt::b200
