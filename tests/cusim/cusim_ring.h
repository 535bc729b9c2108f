// cusim_ring.h — CPU emulation of ring.cuh (mbarrier + cp.async.bulk + dp4a + PDL no-ops).  Test infrastructure; included by
// ring.cuh only when NT_CUSIM is defined (tests/cusim/), never in the product build.
#pragma once
#include "cusim.h"

namespace nt { namespace b200 {

inline uint32_t smem_u32(const void* p) { return (uint32_t)reinterpret_cast<uintptr_t>(p); }
inline void mbar_init(uint64_t* bar, int count) {
    cusim::MBar* b = reinterpret_cast<cusim::MBar*>(bar);
    b->count = (uint16_t)count; b->pending = (uint16_t)count; b->phase = 0; b->tx = 0;
}
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    cusim::MBar* b = reinterpret_cast<cusim::MBar*>(bar);
    if (b->pending == 0) { fprintf(stderr, "cusim: mbarrier over-arrival\n"); cusim::g_failed = true; return; }
    b->tx += (int32_t)bytes;
    b->pending -= 1;
    cusim::mbar_check(b);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    cusim::MBar* b = reinterpret_cast<cusim::MBar*>(bar);
    while (b->phase == (parity & 1u)) {
        if (cusim::g_failed.load()) return;
        cusim::yield("mbarrier wait (TMA data)");
    }
}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15)) {
        fprintf(stderr, "cusim: cp.async.bulk with unaligned dst/src/bytes (%p %p %u)\n", dst, src, bytes);
        cusim::g_failed = true;
    }
    cusim::Cta* c = cusim::g_cta;
    uint8_t* lo = c->dyn_smem.data();
    if (static_cast<uint8_t*>(dst) < lo || static_cast<uint8_t*>(dst) + bytes > lo + c->dyn_smem.size()) {
        fprintf(stderr, "cusim: cp.async.bulk destination outside dynamic shared memory\n");
        cusim::g_failed = true;
        return;
    }
    cusim::PendingCopy pc{dst, src, bytes, bar, 0};
    if (c->copy_delay <= 0) { cusim::complete_copy(pc); return; }
    c->rng = c->rng * 1664525u + 1013904223u;
    pc.delay = 1 + (int)((c->rng >> 16) % (unsigned)c->copy_delay);
    c->copies.push_back(pc);
}
inline int dp4a_us(uint32_t a, int b, int c) {
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 0xFF) * (int)(int8_t)((b >> (8 * i)) & 0xFF);
    return c;
}
inline int dp4a_ss(int a, int b, int c) {
    for (int i = 0; i < 4; i++) c += (int)(int8_t)((a >> (8 * i)) & 0xFF) * (int)(int8_t)((b >> (8 * i)) & 0xFF);
    return c;
}
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}

}}  // namespace nt::b200
