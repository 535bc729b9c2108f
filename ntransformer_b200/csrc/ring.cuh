// ring.cuh — mbarrier + 1-D TMA bulk-copy helpers (sm_100a inline PTX).
// SASS: cp.async.bulk -> UBLKCP.S.G, expect_tx -> SYNCS.ARRIVE.TRANS64, try_wait -> SYNCS.PHASECHK.TRYWAIT.
#pragma once
#ifdef NT_CUSIM            // CPU emulation of this header for tests/cusim (test infrastructure; never defined in the product build)
#include "cusim_ring.h"
#else
#include <cstdint>
#include <cuda_runtime.h>

namespace nt { namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "NT_WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra NT_DONE_%=;\n"
        " bra NT_WAIT_%=;\n"
        "NT_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Global -> shared bulk copy; src, dst and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// u8 x s8 dot product with s32 accumulate (IDP.4A.U8.S8).
__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, int b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_ss(int a_s8x4, int b_s8x4, int c) {
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s8x4), "r"(b_s8x4), "r"(c));
    return d;
}
// Programmatic dependent launch: wait for the producer grid / let dependents start.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}}  // namespace nt::b200
#endif  // NT_CUSIM
