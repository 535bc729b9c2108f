// cusim.h — a small CPU emulator for persistent CUDA kernels (TEST INFRASTRUCTURE, never part of the product path).
//
// The persistent decode kernel (ntransformer_b200/csrc/engine/decode_megakernel.cu) was written while no GPU was
// available.  To find logic errors before hardware time is spent on them, the kernel's own source is compiled here with g++
// (-DNT_CUSIM) and executed on the CPU:
//   * one OS thread per CTA, one ucontext fiber per CUDA thread, cooperative round-robin scheduling inside a CTA;
//   * __syncthreads / __syncwarp / __shfl_*_sync are rendezvous points (a thread that skips one deadlocks the emulation,
//     which the watchdog reports with the place every thread is waiting at — on hardware that would be undefined behaviour);
//   * shared memory: `__shared__` becomes `static thread_local` (one CTA per OS thread), dynamic shared memory is a per-CTA
//     buffer; global memory is ordinary host memory, so several "GPUs" (tensor-parallel ranks) can run in one process and
//     write each other's buffers like NVLink peers;
//   * mbarrier + cp.async.bulk (ring.cuh) are emulated with the same phase/parity/transaction-count rules, the copies
//     themselves complete either immediately or after a random number of scheduler turns (to expose ordering bugs);
//   * acquire/release accesses map to C++ atomics and yield, so spinning threads let the rest of the grid progress.
// What it cannot show: memory-model, proxy-fence and hardware-TMA problems, and anything about speed.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#undef __shared__
#define __shared__ static thread_local
#include <ucontext.h>
#include <sched.h>
#include <time.h>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

namespace cusim {

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    uint3 tid;
    int warp = 0, lane = 0;
    bool done = false;
    const char* where = "running";
    const volatile unsigned* wait_addr = nullptr;   // the scheduler skips the fiber while *wait_addr == wait_val
    unsigned wait_val = 0;
};

struct PendingCopy { void* dst; const void* src; uint32_t bytes; uint64_t* bar; int delay; };

struct Cta {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int cur = 0;
    uint3 bid;
    dim3 gdim, bdim;
    int alive = 0;
    // __syncthreads
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    // warp rendezvous (double-buffered exchange values)
    struct WarpX { uint32_t val[2][32]; int arrived = 0; unsigned gen = 0; };
    std::vector<WarpX> warps;
    std::vector<uint8_t> dyn_smem;
    std::vector<PendingCopy> copies;     // asynchronous bulk copies still "in flight"
    int copy_delay = 0;                  // max scheduler turns a bulk copy stays in flight (0 = synchronous)
    unsigned rng = 12345;
    std::function<void()> body;
};

extern thread_local Cta* g_cta;
extern std::atomic<bool> g_failed;
extern double g_watchdog_s;

inline Fiber& self() { return g_cta->fibers[(size_t)g_cta->cur]; }
inline void yield(const char* where) {
    Fiber& f = self();
    f.where = where;
    swapcontext(&f.ctx, &g_cta->sched);
    f.where = "running";
}
// Block until *addr != val (addr is only written by fibers of the same CTA).
inline void wait_change(const volatile unsigned* addr, unsigned val, const char* where) {
    Fiber& f = self();
    f.wait_addr = addr; f.wait_val = val;
    while (*addr == val) yield(where);
    f.wait_addr = nullptr;
}
inline double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// ---- the emulated mbarrier: [phase:1 | pending arrivals:15 | init count:16] in the low word, pending tx bytes in the high
struct MBar { uint16_t count; uint16_t pending : 15; uint16_t phase : 1; int32_t tx; };
static_assert(sizeof(MBar) == 8, "mbarrier storage is one 64-bit word");
inline void mbar_check(MBar* b) {
    if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->count; }
}
void complete_copy(const PendingCopy& c);
void pump_copies(Cta* c, bool force);

// Runs `body` as a grid of `grid` CTAs x `threads` threads with `dyn_smem` bytes of dynamic shared memory each.  Several
// launches may run concurrently from different host threads (tensor-parallel ranks).  Returns false on deadlock / time-out.
bool launch(int grid, int threads, size_t dyn_smem, int copy_delay, const std::function<void()>& body);

}  // namespace cusim

// ---- built-in variables ---------------------------------------------------------------------------------------------
#define threadIdx (cusim::self().tid)
#define blockIdx (cusim::g_cta->bid)
#define blockDim (cusim::g_cta->bdim)      // NB: include every header that names cudaLaunchConfig_t::gridDim/blockDim BEFORE this one
#define gridDim (cusim::g_cta->gdim)

// ---- synchronisation ---------------------------------------------------------------------------------------------------
inline void __syncthreads() {
    cusim::Cta* c = cusim::g_cta;
    const unsigned gen = c->bar_gen;
    if (++c->bar_arrived >= c->alive) { c->bar_arrived = 0; c->bar_gen++; return; }
    cusim::wait_change(&c->bar_gen, gen, "__syncthreads");
}
inline uint32_t cusim_warp_exchange(uint32_t v, int src_lane_xor, int src_lane_abs) {
    cusim::Cta* c = cusim::g_cta;
    cusim::Fiber& f = cusim::self();
    cusim::Cta::WarpX& w = c->warps[(size_t)f.warp];
    const unsigned gen = w.gen;
    w.val[gen & 1][f.lane] = v;
    if (++w.arrived == 32) { w.arrived = 0; w.gen++; }
    else cusim::wait_change(&w.gen, gen, "warp rendezvous (shuffle / __syncwarp)");
    const int src = src_lane_abs >= 0 ? (src_lane_abs & 31) : (f.lane ^ src_lane_xor);
    return w.val[gen & 1][src];
}
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { cusim_warp_exchange(0, 0, -1); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t b;
    memcpy(&b, &v, 4);
    b = cusim_warp_exchange(b, lane_mask, -1);
    T r;
    memcpy(&r, &b, 4);
    return r;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t b;
    memcpy(&b, &v, 4);
    b = cusim_warp_exchange(b, 0, src_lane);
    T r;
    memcpy(&r, &b, 4);
    return r;
}
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// ---- loads -----------------------------------------------------------------------------------------------------------------
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline T __ldg(const T* p) { return *p; }

// ---- arithmetic intrinsics (IEEE versions of the fast-math forms; compile with -ffp-contract=off) --------------------
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float x) { return expf(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)((src >> (8 * (sel & 7))) & 0xFF);
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;     // sign replication mode
        r |= byte << (8 * i);
    }
    return r;
}
using std::max;
using std::min;
