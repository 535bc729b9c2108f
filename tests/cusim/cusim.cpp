// cusim.cpp — scheduler of the CPU emulator (see cusim.h).  Test infrastructure.
#include "cusim.h"
#include <algorithm>
#include <mutex>

namespace cusim {

thread_local Cta* g_cta = nullptr;
std::atomic<bool> g_failed{false};
double g_watchdog_s = 120.0;
static std::mutex g_print_mu;
static const size_t g_stack_bytes = []() { const char* e = getenv("CUSIM_STACK_KB"); return (size_t)(e ? atoi(e) : 128) * 1024; }();

void complete_copy(const PendingCopy& c) {
    memcpy(c.dst, c.src, c.bytes);
    MBar* b = reinterpret_cast<MBar*>(c.bar);
    b->tx -= (int32_t)c.bytes;
    mbar_check(b);
}
void pump_copies(Cta* c, bool force) {
    for (size_t i = 0; i < c->copies.size();) {
        if (force || --c->copies[i].delay <= 0) {
            complete_copy(c->copies[i]);
            c->copies[i] = c->copies.back();
            c->copies.pop_back();
        } else {
            i++;
        }
    }
}

static void fiber_entry() {
    Cta* c = g_cta;
    c->body();
    Fiber& f = self();
    f.done = true;
    c->alive--;
    if (c->alive > 0 && c->bar_arrived >= c->alive) { c->bar_arrived = 0; c->bar_gen++; }   // exited threads count as arrived
    swapcontext(&f.ctx, &c->sched);
}

static bool run_cta(Cta& c, size_t stack_bytes) {
    g_cta = &c;
    for (size_t i = 0; i < c.fibers.size(); i++) {
        Fiber& f = c.fibers[i];
        f.stack.resize(stack_bytes);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    const double t0 = now_s();
    unsigned pass = 0;
    // CUSIM_SHUFFLE=<seed>: visit the fibers in a fresh pseudo-random order every pass (warps interleave arbitrarily, later
    // threads may run before earlier ones) to expose missing __syncthreads / __syncwarp that round-robin order would hide.
    static const char* shuffle_env = getenv("CUSIM_SHUFFLE");
    unsigned rs = shuffle_env ? (unsigned)atoi(shuffle_env) * 2654435761u + c.bid.x * 40503u + 1u : 0u;
    const size_t nf = c.fibers.size();
    while (c.alive > 0) {
        size_t start = 0, step = 1;
        if (shuffle_env) {                                   // i -> (start + i * step) mod nf with step coprime to nf
            rs = rs * 1664525u + 1013904223u;
            start = (rs >> 8) % nf;
            do { rs = rs * 1664525u + 1013904223u; step = 1 + (rs >> 8) % (nf - 1 ? nf - 1 : 1); } while (std::__gcd(step, nf) != 1);
        }
        for (size_t k = 0; k < nf; k++) {
            const size_t i = (start + k * step) % nf;
            Fiber& f = c.fibers[i];
            if (f.done) continue;
            if (f.wait_addr && *f.wait_addr == f.wait_val) continue;
            c.cur = (int)i;
            swapcontext(&c.sched, &f.ctx);
        }
        pump_copies(&c, false);
        if ((++pass & 63) == 0) {
            sched_yield();
            if (g_failed.load() || now_s() - t0 > g_watchdog_s) {
                std::lock_guard<std::mutex> lk(g_print_mu);
                if (!g_failed.exchange(true)) fprintf(stderr, "cusim: watchdog fired after %.0f s\n", now_s() - t0);
                int shown = 0;
                for (size_t i = 0; i < c.fibers.size() && shown < 6; i++)
                    if (!c.fibers[i].done) {
                        fprintf(stderr, "cusim:   cta %u thread %zu waits at: %s\n", c.bid.x, i, c.fibers[i].where);
                        shown++;
                    }
                return false;
            }
        }
    }
    pump_copies(&c, true);
    return true;
}

bool launch(int grid, int threads, size_t dyn_smem, int copy_delay, const std::function<void()>& body) {
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (int b = 0; b < grid; b++) {
        th.emplace_back([=, &bad, &body]() {
            Cta c;
            c.bid = uint3{(unsigned)b, 0, 0};
            c.gdim = dim3((unsigned)grid);
            c.bdim = dim3((unsigned)threads);
            c.fibers.resize((size_t)threads);
            c.warps.resize((size_t)(threads + 31) / 32);
            c.alive = threads;
            c.dyn_smem.assign(dyn_smem + 128, 0);
            c.copy_delay = copy_delay;
            c.rng = 977u * (unsigned)b + 13u;
            c.body = body;
            for (int t = 0; t < threads; t++) {
                c.fibers[(size_t)t].tid = uint3{(unsigned)t, 0, 0};
                c.fibers[(size_t)t].warp = t / 32;
                c.fibers[(size_t)t].lane = t % 32;
            }
            if (!run_cta(c, g_stack_bytes)) bad++;
        });
    }
    for (auto& t : th) t.join();
    return bad.load() == 0;
}

}  // namespace cusim
